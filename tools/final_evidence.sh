#!/bin/bash
# Round-end evidence on one B200 box: default bench line, reference arm, ncu launch lists and one full capture of the
# threshold and walker kernels.  Everything lands in gpurun_out/<tag>_*; copy what is kept to profiles/.
cd "$(dirname "$0")/.."
tag=${1:-r2w}
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2> gpurun_out/${tag}_ref.err; echo "reference arm rc=$?"
for cfg in C4 C1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 300 --csv --log-file gpurun_out/${tag}_launches_$cfg.csv \
      python bench.py --config $cfg --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-others > gpurun_out/${tag}_launches_$cfg.log 2>&1
  echo "$cfg launches rc=$? lines=$(wc -l < gpurun_out/${tag}_launches_$cfg.csv)"
done
timeout 600 ncu --set full --import-source on --clock-control none -k 'regex:k_threshold_tma|k_trace|k_emit' -s 8 -c 4 -o gpurun_out/${tag}_full \
    python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --skip-others > gpurun_out/${tag}_full.log 2>&1
echo "full capture rc=$?"
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null; echo "raw rows=$(wc -l < gpurun_out/${tag}_full_raw.csv)"
