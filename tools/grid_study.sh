#!/bin/bash
# walker grid study (run on the GPU box): CTAs per SM of the persistent grids of k_trace<false> / k_trace<true>
# (more than 8 CTAs per SM of k_trace<true> need a larger strip budget: ARUCO_B200_TRACE_REC_MB)
run() { ARUCO_B200_GRID_TRACE=$1 ARUCO_B200_GRID_LONG=$2 ARUCO_B200_TRACE_REC_MB=12000 timeout 200 python bench.py --steps 10 --warmup 3 --skip-cpu --skip-e2e --skip-others 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_ms']; print('$1 $2', round(d['value']), round(d['ms_per_step'],3), round(k['trace'],3), round(k['trace_long'],3), d['parity']['ids_exact'])"; }
run 8 8; run 8 4; run 8 6; run 8 12; run 8 16; run 12 8; run 16 8; run 4 4
