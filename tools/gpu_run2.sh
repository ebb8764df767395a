#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2b_tests.txt
tail -4 gpurun_out/r2b_tests.txt
timeout 300 python tools/fuzz_parity.py 40 > gpurun_out/r2b_fuzz.txt 2>&1; tail -3 gpurun_out/r2b_fuzz.txt
timeout 300 python tools/latency.py > gpurun_out/r2b_latency.txt 2>&1; tail -8 gpurun_out/r2b_latency.txt
timeout 120 python tools/octree_timing.py > gpurun_out/r2b_octree_timing.txt 2>&1; head -12 gpurun_out/r2b_octree_timing.txt
for T in 128 256 512; do ORBX_OCTREE_THREADS=$T bash tools/quick_bench.sh r2b_T$T; done
