#!/bin/bash
# resident-path stage times of the default bench (no knn2, no CPU baseline); run on the GPU box: bash tools/quick_bench.sh [tag]
python bench.py --steps 10 --warmup 3 --no-knn2 --no-cpu --no-other 2>/dev/null > gpurun_out/qb_${1:-x}.json
python -c "
import json; d=json.loads(open('gpurun_out/qb_${1:-x}.json').read().strip().splitlines()[-1])
print('${1:-x}', round(d['value']), round(d['ms_per_step'],3), {k:round(v['ms_per_step'],3) for k,v in d['extra']['stages'].items()}, 'e2e', round(d['e2e']['value']))"
