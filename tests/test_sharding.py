"""N>1 paths.  CPU: world_size-2 gloo run of the sharded 2-NN protocol (shard -> local top-2 with global indices ->
all_gather -> lexicographic merge) against the unsharded oracle.  GPU: the CUDA merge kernel on emulated shards."""
import os
import sys

import numpy as np
import pytest

from wut_cuda_orb_slam3_b200 import synth
from wut_cuda_orb_slam3_b200.sharding import merge_top2_reference, shard_rows

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_rows_partition():
    for n in (0, 1, 7, 10, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_rows(n, world, r) for r in range(world)]
            assert sum(c for _, c in spans) == n
            pos = 0
            for first, cnt in spans:
                if cnt:
                    assert first == pos
                    pos += cnt


def _worker(rank, world, port, q_out):
    import torch.distributed as dist
    import torch
    sys.path.insert(0, ROOT)
    from tests import oracle_lib
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    o = oracle_lib.load()
    ndb, nq = 3001, 200
    db = synth.descriptors(5, ndb); q = synth.descriptors(5, nq, is_query=True, ndb=ndb, plant_every=2)
    db[2000] = db[10]; q[0] = db[10]                      # a tie that straddles the shard boundary
    first, cnt = shard_rows(ndb, world, rank)
    idx, dd = o.knn2(q, db[first:first + cnt])
    idx = np.where(idx >= 0, idx + first, -1).astype(np.int32)          # global indices
    ti, td = torch.from_numpy(idx), torch.from_numpy(dd)
    gi = [torch.empty_like(ti) for _ in range(world)]; gd = [torch.empty_like(td) for _ in range(world)]
    dist.all_gather(gi, ti); dist.all_gather(gd, td)
    midx, mdist = merge_top2_reference(torch.stack(gi).numpy(), torch.stack(gd).numpy())
    ridx, rdist = o.knn2(q, db)
    ok = np.array_equal(midx, ridx) and np.array_equal(mdist, rdist) and ridx[0].tolist() == [10, 2000]
    dist.barrier()
    dist.destroy_process_group()
    q_out.put((rank, bool(ok)))


def test_sharded_knn2_protocol_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q_out = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q_out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q_out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
def test_cuda_merge_kernel_on_emulated_shards(oracle):
    import torch
    import wut_cuda_orb_slam3_b200 as orbx
    ndb, nq, S = 50_001, 777, 3
    db = synth.descriptors(9, ndb); q = synth.descriptors(9, nq, is_query=True, ndb=ndb, plant_every=2)
    db[40_000] = db[3]; db[20_000] = db[3]; q[0] = db[3]
    dev = torch.device("cuda:0")
    d_db = torch.from_numpy(db).to(dev); d_q = torch.from_numpy(q).to(dev)
    sh_idx = torch.empty((S, nq, 2), dtype=torch.int32, device=dev); sh_dist = torch.empty_like(sh_idx)
    for s in range(S):
        first, cnt = shard_rows(ndb, S, s)
        orbx.knn2_device(d_q, nq, d_db[first:first + cnt], cnt, sh_idx[s], sh_dist[s], index_base=first)
    out_idx = torch.empty((nq, 2), dtype=torch.int32, device=dev); out_dist = torch.empty_like(out_idx)
    orbx.knn2_merge_device(sh_idx, sh_dist, S, nq, out_idx, out_dist)
    torch.cuda.synchronize()
    ridx, rdist = oracle.knn2(q, db)
    assert np.array_equal(out_idx.cpu().numpy(), ridx) and np.array_equal(out_dist.cpu().numpy(), rdist)
    assert ridx[0].tolist() == [3, 20_000]
    midx, mdist = merge_top2_reference(sh_idx.cpu().numpy(), sh_dist.cpu().numpy())
    assert np.array_equal(midx, ridx) and np.array_equal(mdist, rdist)
