#!/bin/bash
# walker grid study (run on the GPU box): CTAs per SM for k_trace<false> / k_trace<true> / k_emit_long
run() { ARUCO_B200_GRID_TRACE=$1 ARUCO_B200_GRID_LONG=$2 ARUCO_B200_GRID_EMIT=$3 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1 $2 $3', round(d['value']), round(d['roofline']['kernel_ms']['trace'],3), d['parity']['ids_exact'])"; }
run 8 4 8; run 16 4 8; run 8 8 8; run 8 16 8; run 8 4 16; run 16 8 16; run 16 16 16; run 4 4 4; run 12 6 12
