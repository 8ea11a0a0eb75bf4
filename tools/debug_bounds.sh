#!/bin/bash
# Runs the GPU parity / API tests on the AB_DEBUG_BOUNDS build (in-kernel index asserts on the start list, the parked-walk
# and emit queues, the point pool, the quad lists and the walkers' pixel coordinates).  compute-sanitizer is closed on the
# GPU pool; this is the memory-safety check that can run there.  Build first (here, no GPU needed): make debug-bounds
set -e
cd "$(dirname "$0")/.."
export ARUCO_B200_LIB=$PWD/aruco_b200/lib/libaruco_b200_dbg.so
timeout ${1:-900} python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5
python tools/stress_sigma4.py 2>&1 | tail -3
