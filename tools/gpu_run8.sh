#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2g_tests.txt
tail -3 gpurun_out/r2g_tests.txt
bash tools/quick_bench.sh r2g_trim
ORBX_D2H_FULL=1 bash tools/quick_bench.sh r2g_full
bash tools/quick_bench.sh r2g_trim2
