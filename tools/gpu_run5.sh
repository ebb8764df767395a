#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2e_tests.txt
tail -4 gpurun_out/r2e_tests.txt
ORBX_PYR_MULTILEVEL=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2e_tests_ml.txt 2>&1; tail -2 gpurun_out/r2e_tests_ml.txt
ORBX_PYR_MULTILEVEL=0 bash tools/quick_bench.sh r2e_ml0
ORBX_PYR_MULTILEVEL=1 bash tools/quick_bench.sh r2e_ml1
for T in 64 128; do ORBX_OCTREE_THREADS=$T bash tools/quick_bench.sh r2e_T$T; done
for MB in 5 6; do
ORBX_KNN_MINB=$MB timeout 300 python bench.py --steps 3 --warmup 3 --batch 64 --no-cpu --no-other --no-cfg4 > gpurun_out/r2e_knn_$MB.json 2>/dev/null
python -c "
import json; d=json.loads(open('gpurun_out/r2e_knn_$MB.json').read().strip().splitlines()[-1]); k=d['knn2']; print('knn minb $MB', k['value'], k['verified'])"
done
