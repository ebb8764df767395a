#!/bin/bash
# N-GPU bench line under torchrun (N = $1)
N=${1:-4}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $?"; tail -2 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceil', d['e2e'].get('h2d_ceiling_gbs'), 'achieved', d['e2e'].get('h2d_achieved_gbs'))
k=d['knn2']; print('knn2', k['value'], k['verified'], k['nccl_version'])
c=d['cfg4']; print('cfg4', c['frames_per_s'], c['e2e']['value'], c['checksum_matches_n1'])
PY
