#!/bin/bash
# e2e (host buffers, orbx_extract_batch) throughput for several pipeline chunk sizes; run on the GPU box.
for c in 32 64 96 128 256; do
  python bench.py --steps 10 --warmup 3 --no-knn2 --no-cpu --e2e-chunk $c 2>/dev/null > gpurun_out/e2e_c$c.json
  python -c "import json; d=json.loads(open('gpurun_out/e2e_c$c.json').read().strip().splitlines()[-1]); print('chunk', $c, 'resident', round(d['value']), 'e2e', round(d['e2e']['value']))"
done
