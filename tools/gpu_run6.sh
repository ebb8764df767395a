#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2f_tests.txt
tail -4 gpurun_out/r2f_tests.txt
timeout 300 python tools/fuzz_parity.py 40 1234 > gpurun_out/r2f_fuzz.txt 2>&1; tail -2 gpurun_out/r2f_fuzz.txt
timeout 120 python tools/octree_timing.py > gpurun_out/r2f_octree_timing.txt 2>&1; head -12 gpurun_out/r2f_octree_timing.txt
timeout 300 python tools/latency.py > gpurun_out/r2f_latency.txt 2>&1; head -3 gpurun_out/r2f_latency.txt; tail -1 gpurun_out/r2f_latency.txt
ORBX_PYR_MULTILEVEL=1 ORBX_SINGLE_FORK=0 timeout 300 python tools/latency.py > gpurun_out/r2f_latency_ml_nofork.txt 2>&1; head -1 gpurun_out/r2f_latency_ml_nofork.txt; tail -1 gpurun_out/r2f_latency_ml_nofork.txt
bash tools/quick_bench.sh r2f
