#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2c_tests.txt
tail -4 gpurun_out/r2c_tests.txt
timeout 300 python tools/fuzz_parity.py 30 77 > gpurun_out/r2c_fuzz.txt 2>&1; tail -2 gpurun_out/r2c_fuzz.txt
timeout 300 python tools/latency.py > gpurun_out/r2c_latency.txt 2>&1; tail -10 gpurun_out/r2c_latency.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-other --no-cfg4 > gpurun_out/r2c_bench_knn.json 2> gpurun_out/r2c_bench_knn.err; tail -2 gpurun_out/r2c_bench_knn.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2c_bench_knn.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()}, 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])
k=d['knn2']; print('knn2', k['value'], k['verified'], k['roofline']['frac'], k['roofline']['frac_issued'])
PY
