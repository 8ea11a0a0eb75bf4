#!/bin/bash
# ncu launch lists (gpu__time_duration per kernel launch, our kernels only) of every bench config, and the full-set
# capture of the dominant kernel of C4.  Run on the GPU box; results land in gpurun_out/ (copy what is kept to profiles/).
cd "$(dirname "$0")/.."
tag=${1:-r2}
for cfg in C1 C2 C3 C4 C5 C4s4; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/${tag}_launches_$cfg.csv \
      python bench.py --config $cfg --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-others > gpurun_out/${tag}_launches_$cfg.log 2>&1
  echo "$cfg rc=$? lines=$(wc -l < gpurun_out/${tag}_launches_$cfg.csv)"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_threshold_tma -s 6 -c 1 -o gpurun_out/${tag}_thr_full \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-others > gpurun_out/${tag}_thr_full.log 2>&1
echo "full capture rc=$?"
