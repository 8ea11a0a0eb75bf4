#!/bin/bash
# 1 -> N GPU weak-scaling run of bench.py on one box (the driver's SCALE step, runnable by hand): scale_run.sh "8 4 2 1"
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/scale_topo.txt 2>&1
(command -v numactl >/dev/null && numactl -H || lscpu | grep -i numa) > gpurun_out/scale_numa.txt 2>&1
for n in ${1:-8 4 2 1}; do
  if [ "$n" = 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 10 --warmup 3 --skip-others > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale_n$n.json")); e=d["e2e"]
    print("  value %.0f ms/step %.3f per_rank %s" % (d["value"], d["ms_per_step"], [round(x,2) for x in d["per_rank_ms"]]))
    print("  e2e %.0f per_rank_ms %s h2d_only %s frac_of_ceiling %.3f host %s" % (e["value"], [round(x,1) for x in e["per_rank_ms"]], [round(x,1) for x in e["h2d_only_gbs"]], e["frac_of_h2d_ceiling"], d["host"]))
except Exception as ex: print("  parse failed", ex)
PY
done
