for s in 1 2 3 4; do ARUCO_B200_SUBBATCHES=$s timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('sub $s', round(d['value']), d['ms_per_step'], d['parity']['ids_exact'])"; done
