"""Multi-GPU check of orbx_knn2_sharded (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py
Every rank compares the NCCL-merged answer with a single-GPU scan of the whole database and with the CPU oracle, including a
distance tie that straddles a shard boundary and more ranks than rows."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import wut_cuda_orb_slam3_b200 as orbx
from wut_cuda_orb_slam3_b200 import synth
from tests import oracle_lib

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)


def bcast(b):
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if b is not None:
        t.copy_(torch.frombuffer(bytearray(b), dtype=torch.uint8))
    dist.broadcast(t, src=0)
    return bytes(t.cpu().numpy().tobytes())


sm = orbx.ShardedMatcher(rank, world, local, bcast)
o = oracle_lib.load()
ok = True
for (ndb, nq, seed) in [(300_007, 1500, 1), (5, 64, 2), (world - 1 if world > 1 else 1, 10, 3), (64_000, 4097, 4)]:
    db = synth.descriptors(seed, ndb); q = synth.descriptors(seed, nq, is_query=True, ndb=ndb, plant_every=2)
    per = (ndb + world - 1) // world
    if ndb > 2 * per - 1 and per > 3:
        db[per + 1] = db[2]; q[0] = db[2]                       # equal distance in shard 0 and shard 1 -> lower index first
    first, cnt = sm.shard_rows(ndb, world, rank)
    d_db = torch.from_numpy(db[first:first + cnt].copy() if cnt else np.zeros((1, 32), np.uint8)).to(dev)
    d_q = torch.from_numpy(q).to(dev)
    d_idx = torch.full((nq, 2), -9, dtype=torch.int32, device=dev); d_dist = torch.full((nq, 2), -9, dtype=torch.int32, device=dev)
    for _ in range(2):
        sm.knn2(d_q, nq, d_db, cnt, first, d_idx, d_dist)
    torch.cuda.synchronize()
    ridx, rdist = o.knn2(q, db)
    good = np.array_equal(d_idx.cpu().numpy(), ridx) and np.array_equal(d_dist.cpu().numpy(), rdist)
    f_idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); f_dist = torch.empty_like(f_idx)
    orbx.knn2_device(d_q, nq, torch.from_numpy(db).to(dev), ndb, f_idx, f_dist, device=local)
    torch.cuda.synchronize()
    good = good and torch.equal(f_idx, d_idx) and torch.equal(f_dist, d_dist)
    print("rank %d/%d ndb=%d nq=%d: %s" % (rank, world, ndb, nq, "ok" if good else "MISMATCH"), flush=True)
    ok = ok and good
flag = torch.tensor([int(ok)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
sm.close()
dist.destroy_process_group()
if rank == 0:
    print("SHARDED_CHECK", "PASS" if int(flag.item()) else "FAIL", "nccl", flush=True)
sys.exit(0 if int(flag.item()) else 1)
