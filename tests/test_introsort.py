"""The octree kernel's single-lane replay of libstdc++ std::sort (csrc/introsort_replay.h) must reproduce std::sort's
permutation exactly, including among elements the comparator cannot distinguish.  CPU only:
the replay is host+device code; here its host instantiation is compared with the real std::sort in the oracle."""
import numpy as np

from wut_cuda_orb_slam3_b200.capi import lib, ptr


def replay(items):
    a = np.ascontiguousarray(items, np.uint64).copy()
    lib().orbx_debug_sort_replay(ptr(a), len(a))
    return a


def make(keys):
    keys = np.asarray(keys, np.uint64)
    return (keys << np.uint64(24)) | np.arange(len(keys), dtype=np.uint64)   # payload = original index


def test_random_with_many_ties(oracle):
    rng = np.random.default_rng(0)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 217, 434, 1000, 5000]:
        for nkeys in (1, 2, 3, 7, 50, 10 ** 6):
            items = make(rng.integers(0, nkeys, n))
            assert np.array_equal(replay(items), oracle.std_sort_hi40(items)), (n, nkeys)


def test_structured_inputs(oracle):
    for n in (17, 33, 128, 1025, 4096):
        for keys in (np.arange(n), np.arange(n)[::-1], np.zeros(n), np.arange(n) % 2, np.arange(n) // 3,
                     np.concatenate([np.arange(n // 2), np.arange(n - n // 2)])):
            items = make(keys)
            assert np.array_equal(replay(items), oracle.std_sort_hi40(items))


def test_median_of_three_killer_hits_heapsort_fallback(oracle):
    # classic anti-quicksort sequence for median-of-3 pivots: forces the depth limit -> heap-sort fallback
    n = 4096
    k = n // 2
    keys = np.zeros(n, np.int64)
    for i in range(1, k + 1):
        if i % 2 == 1:
            keys[i - 1] = i
            keys[i] = k + i
        keys[k + i - 1] = 2 * i
    items = make(keys)
    assert np.array_equal(replay(items), oracle.std_sort_hi40(items))


def test_octree_like_keys(oracle):
    rng = np.random.default_rng(1)
    for n in (20, 60, 150, 400):
        cnt = rng.integers(2, 12, n).astype(np.uint64)
        ulx = (rng.integers(0, 8, n) * 45).astype(np.uint64)
        items = make((cnt << np.uint64(13)) | ulx)
        assert np.array_equal(replay(items), oracle.std_sort_hi40(items))


def test_32bit_rank_items_give_the_same_permutation(oracle):
    """The kernel sorts (dense rank << 16 | creation index) when there are <= 512 nodes; same permutation as the 64-bit items."""
    rng = np.random.default_rng(7)
    for n in (1, 2, 17, 40, 118, 300, 512):
        for nkeys in (1, 3, 20, 10 ** 6):
            keys = rng.integers(0, nkeys, n).astype(np.uint64)
            ref = oracle.std_sort_hi40(make(keys))
            rank = np.array([(keys < k).sum() for k in keys], np.uint32)
            items = np.ascontiguousarray((rank << np.uint32(16)) | np.arange(n, dtype=np.uint32))
            lib().orbx_debug_sort_replay32(ptr(items), n)
            assert np.array_equal(items & np.uint32(0xffff), (ref & np.uint64(0xffffff)).astype(np.uint32)), (n, nkeys)


def replay_ranges(items):
    a = np.ascontiguousarray(items, np.uint64).copy()
    lib().orbx_debug_sort_replay_ranges(ptr(a), len(a))
    return a


def test_range_parallel_formulation_matches_std_sort(oracle):
    """The octree kernel runs the introsort as rounds of independent range partitions + a block-wise stable placement
    (csrc/introsort_replay.h, range-parallel formulation); its host simulation must give std::sort's permutation."""
    rng = np.random.default_rng(3)
    for n in list(range(0, 40)) + [63, 64, 65, 100, 118, 217, 434, 1000, 4000]:
        for nkeys in (1, 2, 3, 7, 50, 10 ** 6):
            items = make(rng.integers(0, nkeys, n))
            assert np.array_equal(replay_ranges(items), oracle.std_sort_hi40(items)), (n, nkeys)
    for n in (17, 33, 128, 1025, 4096):
        for keys in (np.arange(n), np.arange(n)[::-1], np.zeros(n), np.arange(n) % 2, np.arange(n) // 3,
                     np.concatenate([np.arange(n // 2), np.arange(n - n // 2)])):
            items = make(keys)
            assert np.array_equal(replay_ranges(items), oracle.std_sort_hi40(items))
    for n in (20, 60, 150, 400):
        cnt = rng.integers(2, 12, n).astype(np.uint64)
        ulx = (rng.integers(0, 8, n) * 45).astype(np.uint64)
        items = make((cnt << np.uint64(13)) | ulx)
        assert np.array_equal(replay_ranges(items), oracle.std_sort_hi40(items))


def test_range_parallel_heapsort_fallback(oracle):
    n = 4096
    k = n // 2
    keys = np.zeros(n, np.int64)
    for i in range(1, k + 1):
        if i % 2 == 1:
            keys[i - 1] = i
            keys[i] = k + i
        keys[k + i - 1] = 2 * i
    items = make(keys)
    assert np.array_equal(replay_ranges(items), oracle.std_sort_hi40(items))


def test_range_parallel_32bit_items(oracle):
    rng = np.random.default_rng(9)
    for n in (1, 2, 17, 40, 118, 300, 512):
        for nkeys in (1, 3, 20, 10 ** 6):
            keys = rng.integers(0, nkeys, n).astype(np.uint64)
            ref = oracle.std_sort_hi40(make(keys))
            rank = np.array([(keys < k).sum() for k in keys], np.uint32)
            items = np.ascontiguousarray((rank << np.uint32(16)) | np.arange(n, dtype=np.uint32))
            lib().orbx_debug_sort_replay_ranges32(ptr(items), n)
            assert np.array_equal(items & np.uint32(0xffff), (ref & np.uint64(0xffffff)).astype(np.uint32)), (n, nkeys)


def replay_ranked(items):
    a = np.ascontiguousarray(items, np.uint64).copy()
    lib().orbx_debug_sort_replay_ranked.restype = None
    lib().orbx_debug_sort_replay_ranked(ptr(a), len(a))
    return a


def test_rank_arithmetic_partition_matches_std_sort(oracle):
    """The Hoare partition as rank arithmetic (k-th stop of the left scan pairs with the k-th stop of the right scan while they
    have not crossed; csrc/introsort_replay.h unguarded_partition_ranked_) — the form a warp evaluates cooperatively in the
    octree kernel — must give std::sort's permutation on everything the scan / swap loop is tested on."""
    rng = np.random.default_rng(11)
    for n in list(range(0, 48)) + [63, 64, 65, 100, 118, 217, 300, 434, 512, 1000, 4000]:
        for nkeys in (1, 2, 3, 7, 50, 10 ** 6):
            for rep in range(3):
                items = make(rng.integers(0, nkeys, n))
                assert np.array_equal(replay_ranked(items), oracle.std_sort_hi40(items)), (n, nkeys)
    for n in (17, 33, 128, 1025, 4096):
        for keys in (np.arange(n), np.arange(n)[::-1], np.zeros(n), np.arange(n) % 2, np.arange(n) // 3,
                     np.concatenate([np.arange(n // 2), np.arange(n - n // 2)])):
            items = make(keys)
            assert np.array_equal(replay_ranked(items), oracle.std_sort_hi40(items))
    for n in (20, 60, 150, 400):
        for rep in range(20):
            cnt = rng.integers(2, 12, n).astype(np.uint64)
            ulx = (rng.integers(0, 8, n) * 45).astype(np.uint64)
            items = make((cnt << np.uint64(13)) | ulx)
            assert np.array_equal(replay_ranked(items), oracle.std_sort_hi40(items))
    # median-of-three killer: depth limit -> heap-sort fallback on the way
    n = 4096
    k = n // 2
    keys = np.zeros(n, np.int64)
    for i in range(1, k + 1):
        if i % 2 == 1:
            keys[i - 1] = i
            keys[i] = k + i
        keys[k + i - 1] = 2 * i
    items = make(keys)
    assert np.array_equal(replay_ranked(items), oracle.std_sort_hi40(items))
    # 32-bit rank items
    lib().orbx_debug_sort_replay_ranked32.restype = None
    for n in (1, 2, 17, 40, 118, 300, 512):
        for nkeys in (1, 3, 20, 10 ** 6):
            keys = rng.integers(0, nkeys, n).astype(np.uint64)
            ref = oracle.std_sort_hi40(make(keys))
            rank = np.array([(keys < k).sum() for k in keys], np.uint32)
            items = np.ascontiguousarray((rank << np.uint32(16)) | np.arange(n, dtype=np.uint32))
            lib().orbx_debug_sort_replay_ranked32(ptr(items), n)
            assert np.array_equal(items & np.uint32(0xffff), (ref & np.uint64(0xffffff)).astype(np.uint32)), (n, nkeys)
