#!/bin/bash
# same-box A/B of chunk schedules of orbx_extract_batch: default (ramp up + ramp down) against ORBX_BATCH_SCHED overrides
run() { python bench.py --steps 6 --warmup 2 --no-knn2 --no-cpu --no-other --no-cfg4 --e2e-batch 8192 --e2e-chunk ${CH:-256} 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', 'e2e', round(d['e2e']['value']), 'ceiling GB/s', round(d['e2e']['h2d_ceiling_gbs'], 1))"; }
for rep in 1 2; do
  run default
  ORBX_BATCH_SCHED="64,128,256" run old_rampup_only
  ORBX_BATCH_SCHED="32,64,128,256" run finer_rampup
done
