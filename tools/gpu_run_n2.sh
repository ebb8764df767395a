#!/bin/bash
# 2-GPU pass: sharded 2-NN through the C ABI over NCCL (small cases against the oracle), then the bench at N=2
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py > gpurun_out/r2_sharded_check_n2.txt 2>&1; echo "sharded_check exit $?"; tail -10 gpurun_out/r2_sharded_check_n2.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench exit $?"; tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n2.json').read().strip().splitlines()[-1])
print('N=2', round(d['value']), 'e2e', round(d['e2e']['value']), 'ceil', d['e2e'].get('h2d_ceiling_gbs'))
k=d['knn2']; print('knn2', k['value'], k['verified'], k['verification'], k['nccl_version'])
c=d['cfg4']; print('cfg4', c['frames_per_s'], c['e2e']['value'], c['checksum_matches_n1'], c['checksum'])
PY
