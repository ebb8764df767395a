#!/bin/bash
# e2e (host buffers, orbx_extract_batch) throughput for several (frames per call, pipeline chunk) pairs; run on the GPU box.
for cfg in ${E2E_CFGS:-4096:256 8192:256 8192:384 8192:512 16384:512}; do
  set -- ${cfg/:/ }
  python bench.py --steps 4 --warmup 2 --no-knn2 --no-cpu --no-other --no-cfg4 --e2e-batch $1 --e2e-chunk $2 2>/dev/null > gpurun_out/e2e_sweep.json
  python -c "import json; d=json.loads(open('gpurun_out/e2e_sweep.json').read().strip().splitlines()[-1]); print('frames/call', $1, 'chunk', $2, 'resident', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
