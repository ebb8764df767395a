#!/bin/bash
# round-2 first GPU pass: tests, latency, octree phase timing, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.txt 2>&1; echo "pytest exit $?" >> gpurun_out/r2a_tests.txt
tail -5 gpurun_out/r2a_tests.txt
timeout 300 python tools/latency.py > gpurun_out/r2a_latency.txt 2>&1; tail -8 gpurun_out/r2a_latency.txt
ORBX_SINGLE_FORK=0 timeout 300 python tools/latency.py > gpurun_out/r2a_latency_nofork.txt 2>&1; head -3 gpurun_out/r2a_latency_nofork.txt
timeout 120 python tools/octree_timing.py > gpurun_out/r2a_octree_timing.txt 2>&1; head -12 gpurun_out/r2a_octree_timing.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench exit $?"; tail -3 gpurun_out/r2a_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2a_bench.json').read().strip().splitlines()[-1])
    print(round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()}, 'e2e', round(d['e2e']['value']), d['e2e'].get('h2d_ceiling_gbs'))
    print('knn2', d.get('knn2',{}).get('value'), d.get('knn2',{}).get('verified'), d.get('knn2',{}).get('verification'))
    print('cfg4', {k:v for k,v in d.get('cfg4',{}).items() if k in ('frames_per_s','ms_per_pass','checksum','e2e')})
    print('other', json.dumps(d['extra'].get('other_configs'))[:600])
except Exception as e:
    print('parse failed', e)
PY
