"""Host-side sharding rules for the two multi-GPU workloads.

* frames: independent units -> contiguous static partition, no collective on the data path;
* 2-NN database: queries replicated, database rows split into contiguous shards, every rank returns (d1,i1,d2,i2) with
  GLOBAL row indices, one all-gather of nq x 16 B per rank, then a lexicographic (distance, index) merge
  (orbx_knn2_merge_device) that reproduces the single-GPU / BFMatcher tie rule exactly.
"""


def shard_rows(n_rows, world, rank):
    """Contiguous shard [first, first + count) of n_rows for `rank` of `world` (last shards may be empty)."""
    per = (n_rows + world - 1) // world
    first = min(rank * per, n_rows)
    return first, max(0, min(per, n_rows - first))


def merge_top2_reference(idx_shards, dist_shards):
    """Pure-numpy statement of the merge rule (used by tests; the product path is the CUDA merge kernel)."""
    import numpy as np
    S, nq, _ = idx_shards.shape
    idx = np.full((nq, 2), -1, np.int32)
    dist = np.full((nq, 2), np.iinfo(np.int32).max, np.int32)
    for q in range(nq):
        cands = sorted((int(dist_shards[s, q, k]), int(idx_shards[s, q, k])) for s in range(S) for k in range(2) if idx_shards[s, q, k] >= 0)
        for k, (d, i) in enumerate(cands[:2]):
            dist[q, k], idx[q, k] = d, i
    return idx, dist
