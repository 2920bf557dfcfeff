// merge.cuh — K4: merge G per-shard best-first result lists into one (multi-GPU top-k merge,
// runs after the NCCL gather).  Replaces faiss's per-shard heap merge [upstream IndexShards /
// HeapArray merge]; there is no analogue in memo_cli.py (single process).
//
// Rank-by-counting merge: every candidate (g,j) computes its final rank as the number of
// candidates that precede it under (score best-first, lower shard first, earlier position first)
// with one binary search per shard list, and writes itself to out[rank] if rank < k.  Ranks are a
// permutation of 0..G*k-1, so every output slot is written exactly once — no shared memory, any k.
#pragma once
#include "common.cuh"

// Padding entries (a shard with fewer than k candidates) carry the sentinel score -/+FLT_MAX, which no candidate can
// have; they are recognised by it, not by a negative id — negative record ids are legal in an id map.
template <int METRIC>
__device__ __forceinline__ uint32_t merge_hi(float s, int64_t id) {
    (void)id;
    return b200_score_valid<METRIC>(s) ? b200_key_hi<METRIC>(s) : 0u;
}

// lists are descending in hi.  number of entries with hi > h (strict) or hi >= h
template <int METRIC>
__device__ __forceinline__ int64_t count_better(const float* D, const int64_t* I, int64_t k,
                                                uint32_t h, bool or_equal) {
    int64_t lo = 0, hi = k;  // first index where the predicate fails
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        uint32_t hm = merge_hi<METRIC>(D[mid], I[mid]);
        bool pred = or_equal ? (hm >= h) : (hm > h);
        if (pred) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <int METRIC>
__global__ void __launch_bounds__(256)
merge_topk_kernel(int G, int64_t nq, int64_t k, const float* __restrict__ Dp,
                  const int64_t* __restrict__ Ip, int64_t dstride, int64_t istride,
                  float* __restrict__ Do, int64_t* __restrict__ Io) {
    const int64_t total = (int64_t)G * nq * k;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        // t enumerates (q, g, j)
        int64_t q = t / (G * k);
        int64_t r = t - q * (G * k);
        int g = (int)(r / k);
        int64_t j = r - (int64_t)g * k;
        const float* Dme = Dp + (int64_t)g * dstride + q * k;
        const int64_t* Ime = Ip + (int64_t)g * istride + q * k;
        float s = Dme[j];
        int64_t id = Ime[j];
        uint32_t h = merge_hi<METRIC>(s, id);
        int64_t rank = j;
        for (int g2 = 0; g2 < G; ++g2) {
            if (g2 == g) continue;
            const float* D2 = Dp + (int64_t)g2 * dstride + q * k;
            const int64_t* I2 = Ip + (int64_t)g2 * istride + q * k;
            rank += count_better<METRIC>(D2, I2, k, h, /*or_equal=*/g2 < g);
        }
        if (rank < k) {
            Do[q * k + rank] = h == 0u ? ((METRIC == 0) ? -FLT_MAX : FLT_MAX) : s;
            Io[q * k + rank] = h == 0u ? (int64_t)-1 : id;
        }
    }
}

// Certificate of a merged row-sharded batch (K3 per shard, cabi.cu b200_index_search_shard_dev): query q is proven
// exact iff its merged k-th entry (k = min(k, rows in all shards)) exists and strictly beats every shard's bound —
// nothing outside the shards' lists can then belong to the top k.  uncertified[q] = 1 and *n_uncertified counts the rest.
template <int METRIC>
__global__ void __launch_bounds__(256)
merge_certify_kernel(int G, int64_t nq, int64_t k, int64_t want, const float* __restrict__ Do,
                     const float* __restrict__ bounds, int64_t bstride, int* __restrict__ uncertified,
                     int* __restrict__ n_uncertified) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    bool ok = true;
    if (want > 0) {
        const float kth = Do[q * k + want - 1];
        ok = b200_score_valid<METRIC>(kth);
        for (int g = 0; g < G && ok; ++g) {
            const float b = bounds[(int64_t)g * bstride + q];
            ok = METRIC == 0 ? (kth > b) : (kth < b);
        }
    }
    uncertified[q] = ok ? 0 : 1;
    if (!ok) atomicAdd(n_uncertified, 1);
}

// bound that excludes nothing: the list it travels with is this shard's exact top k
template <int METRIC>
__global__ void fill_neutral_bound_kernel(float* bound, int64_t nq) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nq) bound[q] = METRIC == 0 ? -INFINITY : INFINITY;
}
