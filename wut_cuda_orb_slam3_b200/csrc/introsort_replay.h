// introsort_replay.h — exact replay of libstdc++'s std::sort (GCC 13 bits/stl_algo.h, bits/stl_heap.h) on an array
// of 64-bit items compared by their high 40 bits only (the low 24 bits are payload carried along).
//
// Why: DistributeOctTree sorts its expandable nodes with std::sort(compareNodes) (reference
// src/ORBextractor.cc:569-582, 709), whose comparator orders only by (count, UL.x).  Ties are broken by whatever
// permutation introsort happens to produce, and that permutation decides which nodes are split before the
// `size >= N` break and the order children are pushed — i.e. the keypoint set and order.
// So the GPU octree replays the algorithm step for step: introsort loop with threshold 16 and depth limit
// 2*floor(log2 n), median-of-three of (first+1, mid, last-1) moved to first, unguarded Hoare partition, heap-sort
// fallback, then the final insertion sort (guarded for the first 16, unguarded afterwards).
// tests/test_introsort.py checks the replay against std::sort.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ORBX_SORT_HD __host__ __device__ inline
#else
#define ORBX_SORT_HD inline
#endif

namespace orbx_sort {

constexpr int kPayloadBits = 24;
ORBX_SORT_HD unsigned long long make_item(unsigned long long key, uint32_t payload) { return (key << kPayloadBits) | (payload & 0xffffffu); }
ORBX_SORT_HD uint32_t payload(unsigned long long v) { return (uint32_t)(v & 0xffffffu); }
// 64-bit items: compared by the bits above the 24-bit payload
ORBX_SORT_HD bool lt(unsigned long long a, unsigned long long b) { return (a >> kPayloadBits) < (b >> kPayloadBits); }
// 32-bit items: rank << 16 | payload, compared by rank only:  rank(a) < rank(b)  <=>  (a | 0xffff) < b
ORBX_SORT_HD bool lt(uint32_t a, uint32_t b) { return (a | 0xffffu) < b; }
template <typename item_t> ORBX_SORT_HD void swp(item_t* a, item_t* b) { item_t t = *a; *a = *b; *b = t; }

template <typename item_t> ORBX_SORT_HD void push_heap_(item_t* first, int hole, int top, item_t value)
{
    int parent = (hole - 1) / 2;
    while (hole > top && lt(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

template <typename item_t> ORBX_SORT_HD void adjust_heap_(item_t* first, int hole, int len, item_t value)
{
    const int top = hole;
    int second = hole;
    while (second < (len - 1) / 2) {
        second = 2 * (second + 1);
        if (lt(first[second], first[second - 1])) second--;
        first[hole] = first[second];
        hole = second;
    }
    if ((len & 1) == 0 && second == (len - 2) / 2) {
        second = 2 * (second + 1);
        first[hole] = first[second - 1];
        hole = second - 1;
    }
    push_heap_(first, hole, top, value);
}

// std::__partial_sort(first, last, last) == make_heap + sort_heap
template <typename item_t> ORBX_SORT_HD void heap_sort_(item_t* first, int len)
{
    if (len >= 2) {
        int parent = (len - 2) / 2;
        while (true) {
            item_t v = first[parent];
            adjust_heap_(first, parent, len, v);
            if (parent == 0) break;
            parent--;
        }
    }
    int last = len;
    while (last > 1) {
        --last;
        item_t v = first[last];
        first[last] = first[0];
        adjust_heap_(first, 0, last, v);
    }
}

template <typename item_t> ORBX_SORT_HD void move_median_to_first_(item_t* result, item_t* a, item_t* b, item_t* c)
{
    if (lt(*a, *b)) {
        if (lt(*b, *c)) swp(result, b);
        else if (lt(*a, *c)) swp(result, c);
        else swp(result, a);
    } else if (lt(*a, *c)) swp(result, a);
    else if (lt(*b, *c)) swp(result, c);
    else swp(result, b);
}

// The scans below look 4 elements ahead: the probes of one step are independent loads, so a single GPU lane is not
// serialised on one shared-memory round trip per comparison.  `n` bounds the speculative reads; decisions are taken in
// exactly the order of the libstdc++ loops.
template <typename item_t> ORBX_SORT_HD int unguarded_partition_(item_t* base, int n, int first, int last, int pivot)
{
    const item_t pv = base[pivot];
    while (true) {
        while (true) {                                        // while (comp(first, pivot)) ++first;
            const item_t a0 = base[first], a1 = base[first + 1 < n ? first + 1 : n - 1],
                         a2 = base[first + 2 < n ? first + 2 : n - 1], a3 = base[first + 3 < n ? first + 3 : n - 1];
            if (!lt(a0, pv)) break;
            ++first;
            if (!lt(a1, pv)) break;
            ++first;
            if (!lt(a2, pv)) break;
            ++first;
            if (!lt(a3, pv)) break;
            ++first;
        }
        --last;
        while (true) {                                        // while (comp(pivot, last)) --last;
            const item_t a0 = base[last], a1 = base[last - 1 > 0 ? last - 1 : 0], a2 = base[last - 2 > 0 ? last - 2 : 0],
                         a3 = base[last - 3 > 0 ? last - 3 : 0];
            if (!lt(pv, a0)) break;
            --last;
            if (!lt(pv, a1)) break;
            --last;
            if (!lt(pv, a2)) break;
            --last;
            if (!lt(pv, a3)) break;
            --last;
        }
        if (!(first < last)) return first;
        swp(base + first, base + last);
        ++first;
    }
}

// The same outcome from rank arithmetic instead of the scan / swap loop (the form a warp evaluates cooperatively, see
// sort_replay_parallel in kernels_octree.cu).  Let L_0 < L_1 < ... be the positions in [first, last) whose element does NOT
// compare less than the pivot (where the left scan can stop) and R_0 > R_1 > ... those whose element the pivot does NOT
// compare less than (where the right scan can stop).  Between two swaps the scans only cross untouched elements, so swap k
// exchanges L_k and R_k as long as L_k < R_k; with K such swaps the final left scan stops at the first CURRENT element that
// is not less than the pivot to the right of L_{K-1}: the untouched L_K, or R_{K-1} (which now holds the old a[L_{K-1}]),
// whichever comes first.  posL / posR: scratch for last - first positions each.
template <typename item_t> ORBX_SORT_HD int unguarded_partition_ranked_(item_t* base, int first, int last, int pivot, int* posL, int* posR)
{
    const item_t pv = base[pivot];
    int nl = 0, nr = 0;
    for (int i = first; i < last; ++i)
        if (!lt(base[i], pv)) posL[nl++] = i;
    for (int i = last - 1; i >= first; --i)
        if (!lt(pv, base[i])) posR[nr++] = i;
    int K = 0;
    while (K < nl && K < nr && posL[K] < posR[K]) ++K;
    for (int k = 0; k < K; ++k) swp(base + posL[k], base + posR[k]);
    int ret = last;                                     // (the median-of-three pivot guarantees a stop inside the range)
    if (K < nl) ret = posL[K];
    if (K > 0 && posR[K - 1] < ret) ret = posR[K - 1];
    return ret;
}

template <typename item_t> ORBX_SORT_HD void unguarded_linear_insert_(item_t* base, int last)
{
    const item_t val = base[last];
    int next = last - 1;
    while (true) {
        // positions <= next are never written by the shifts of this step, so they can be read ahead
        const item_t a0 = base[next > 0 ? next : 0], a1 = base[next - 1 > 0 ? next - 1 : 0], a2 = base[next - 2 > 0 ? next - 2 : 0],
                     a3 = base[next - 3 > 0 ? next - 3 : 0];
        if (!lt(val, a0)) break;
        base[last] = a0; last = next; --next;
        if (!lt(val, a1)) break;
        base[last] = a1; last = next; --next;
        if (!lt(val, a2)) break;
        base[last] = a2; last = next; --next;
        if (!lt(val, a3)) break;
        base[last] = a3; last = next; --next;
    }
    base[last] = val;
}

template <typename item_t> ORBX_SORT_HD void insertion_sort_(item_t* base, int first, int last)
{
    if (first == last) return;
    for (int i = first + 1; i != last; ++i) {
        if (lt(base[i], base[first])) {
            item_t val = base[i];
            for (int j = i; j > first; --j) base[j] = base[j - 1];
            base[first] = val;
        } else {
            unguarded_linear_insert_(base, i);
        }
    }
}

// std::sort(base, base + n, comp) with comp = "high 32 bits less-than".
template <typename item_t> ORBX_SORT_HD void sort_replay(item_t* base, int n)
{
    if (n <= 0) return;
    // __introsort_loop with an explicit stack (the recursion is on the right part, the loop on the left part)
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    struct Frame { int first, last, depth; };
    Frame stack[72];
    int sp = 0;
    stack[sp++] = Frame{0, n, 2 * lg};
    while (sp > 0) {
        Frame f = stack[--sp];
        int first = f.first, last = f.last, depth = f.depth;
        while (last - first > 16) {
            if (depth == 0) {
                heap_sort_(base + first, last - first);
                break;
            }
            --depth;
            const int mid = first + (last - first) / 2;
            move_median_to_first_(base + first, base + first + 1, base + mid, base + last - 1);
            const int cut = unguarded_partition_(base, n, first + 1, last, first);
            // recursion: __introsort_loop(cut, last, depth) runs BEFORE the loop continues on [first, cut); the two
            // ranges are disjoint so the order of processing does not change the result.
            if (sp < 72) stack[sp++] = Frame{cut, last, depth};
            last = cut;
        }
    }
    // __final_insertion_sort
    if (n > 16) {
        insertion_sort_(base, 0, 16);
        for (int i = 16; i != n; ++i) unguarded_linear_insert_(base, i);
    } else {
        insertion_sort_(base, 0, n);
    }
}

// ---- range-parallel formulation ----------------------------------------------------------------------------------------
// The introsort loop recurses on disjoint ranges, so the partition steps of one recursion level are independent: every
// range longer than 16 is partitioned by its own lane (the partition itself stays the exact sequential Hoare scan), the
// children go to the next round.  __final_insertion_sort is a STABLE sort of each leftover block of <= 16 elements: after
// the partitions every element left of a block compares <= and every element right of it >=, so an insertion never leaves
// its block (a guarded insertion into block 0 equals an unguarded one that stops at index 0), and a stable sort of a block
// is unique — it is evaluated by rank counting, one lane per element.  Ranges whose depth budget is exhausted are
// heap-sorted by their lane and left alone (the insertion pass finds them already ordered).
// Same permutation as sort_replay() / std::sort; ~2n dependent steps instead of ~n log n.
struct Range { int first, last, depth; };

// One introsort step on r (last - first > 16).  Returns the number of children written to out (0: heap-sorted).
template <typename item_t> ORBX_SORT_HD int range_step(item_t* base, int n, Range r, Range out[2])
{
    if (r.depth == 0) {
        heap_sort_(base + r.first, r.last - r.first);
        return 0;
    }
    const int depth = r.depth - 1;
    const int mid = r.first + (r.last - r.first) / 2;
    move_median_to_first_(base + r.first, base + r.first + 1, base + mid, base + r.last - 1);
    const int cut = unguarded_partition_(base, n, r.first + 1, r.last, r.first);
    out[0] = Range{r.first, cut, depth};
    out[1] = Range{cut, r.last, depth};
    return 2;
}

// range_step with the ranked partition (host check of the formula the device evaluates per warp)
template <typename item_t> inline int range_step_ranked(item_t* base, int n, Range r, Range out[2], int* posL, int* posR)
{
    if (r.depth == 0) {
        heap_sort_(base + r.first, r.last - r.first);
        return 0;
    }
    const int depth = r.depth - 1;
    const int mid = r.first + (r.last - r.first) / 2;
    move_median_to_first_(base + r.first, base + r.first + 1, base + mid, base + r.last - 1);
    const int cut = unguarded_partition_ranked_(base, r.first + 1, r.last, r.first, posL, posR);
    out[0] = Range{r.first, cut, depth};
    out[1] = Range{cut, r.last, depth};
    return 2;
}

// Position of element i after the stable sort of its block [f, l).
template <typename item_t> ORBX_SORT_HD int block_stable_pos(const item_t* base, int f, int l, int i)
{
    const item_t e = base[i];
    int rank = 0;
    for (int j = f; j < l; ++j) rank += (lt(base[j], e) || (j < i && !lt(e, base[j]))) ? 1 : 0;
    return f + rank;
}

ORBX_SORT_HD int initial_depth(int n)
{
    int lg = 0;
    for (int t = n; t > 1; t >>= 1) ++lg;
    return 2 * lg;
}

// Host-side sequential simulation of the range-parallel schedule (rounds of independent range steps, then the block-wise
// stable placement): the device version in kernels_octree.cu runs the same steps with one lane per range / element.
// blk[n]: packed block bounds (f | l << 16) per element; tmp[n]: output buffer.  n < 65536.
template <typename item_t> inline void sort_replay_ranges_host(item_t* base, int n, uint32_t* blk, item_t* tmp, bool ranked = false)
{
    if (n <= 0) return;
    int* posL = ranked ? new int[2 * (size_t)n] : nullptr;
    int* posR = ranked ? posL + n : nullptr;
    Range cur[256], nxt[256];          // ranges longer than 16 elements: at most n / 17 of them
    int nc = 0, nn = 0;
    auto settle = [&](Range r, bool sorted) {
        for (int i = r.first; i < r.last; ++i) blk[i] = sorted ? ((uint32_t)i | ((uint32_t)(i + 1) << 16)) : ((uint32_t)r.first | ((uint32_t)r.last << 16));
    };
    Range all{0, n, initial_depth(n)};
    if (n > 16) cur[nc++] = all; else settle(all, false);
    while (nc > 0) {
        nn = 0;
        for (int q = 0; q < nc; ++q) {
            Range out[2];
            const int k = ranked ? range_step_ranked(base, n, cur[q], out, posL, posR) : range_step(base, n, cur[q], out);
            if (k == 0) settle(cur[q], true);
            for (int c = 0; c < k; ++c) {
                if (out[c].last - out[c].first > 16) nxt[nn++] = out[c];
                else settle(out[c], false);
            }
        }
        for (int q = 0; q < nn; ++q) cur[q] = nxt[q];
        nc = nn;
    }
    for (int i = 0; i < n; ++i) tmp[block_stable_pos(base, (int)(blk[i] & 0xffffu), (int)(blk[i] >> 16), i)] = base[i];
    for (int i = 0; i < n; ++i) base[i] = tmp[i];
    delete[] posL;
}

}  // namespace orbx_sort
