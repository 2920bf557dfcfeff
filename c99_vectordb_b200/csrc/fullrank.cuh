// fullrank.cuh — the general-k path: a full ranking of the database for one query.
//
// memo never asks for a small k: search_all() calls index.search(q, k = ntotal)
// (memo_cli.py:291-292) and post-filters in Python, so the drop-in must return a complete
// best-first ranking.  For k above the fused limit the scan kernel writes one 32-bit score key
// per row (4 B/row next to the d*4 B/row it reads) and this file sorts them: a stable LSD radix
// sort (4 passes x 8 bits) on the inverted key carrying the row as payload.  Stability gives the
// tie rule for free: rows start in ascending order, so equal scores stay smaller-row-first.
#pragma once
#include "common.cuh"

#define RADIX_THREADS 256
#define RADIX_ITEMS 16
#define RADIX_CHUNK (RADIX_THREADS * RADIX_ITEMS)

// inv = ~hi so that ascending inv == best-first; invalid rows (hi == 0) sort last
__global__ void __launch_bounds__(256) fullrank_prepare_kernel(const uint32_t* hi_keys,
                                                               uint32_t* inv_keys, uint32_t* rows,
                                                               uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        inv_keys[i] = ~hi_keys[i];
        rows[i] = (uint32_t)i;
    }
}

// hist[bin * nblocks + block] = number of keys of this block's chunk with that digit
__global__ void __launch_bounds__(RADIX_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift,
                  uint32_t* __restrict__ hist, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RADIX_CHUNK;
    for (int r = 0; r < RADIX_ITEMS; ++r) {
        uint64_t i = base + (uint64_t)r * RADIX_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(size_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

// in-place exclusive scan of `count` uint32 by one CTA (count = 256 * nblocks)
__global__ void __launch_bounds__(1024) radix_scan_kernel(uint32_t* data, uint64_t count) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < count; base += 1024) {
        uint64_t i = base + threadIdx.x;
        uint32_t v = i < count ? data[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(B200_FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t w = warp_tot[lane];
            uint32_t xs = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t y = __shfl_up_sync(B200_FULL_MASK, xs, o);
                if (lane >= o) xs += y;
            }
            warp_tot[lane] = xs - w;  // exclusive warp offsets
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t incl = x + warp_tot[warp] + carry;
        if (i < count) data[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = incl;
        __syncthreads();
    }
}

// stable scatter of one chunk per CTA
__global__ void __launch_bounds__(RADIX_THREADS)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                     uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out, uint64_t n,
                     int shift, const uint32_t* __restrict__ offsets, uint32_t nblocks) {
    constexpr int NW = RADIX_THREADS / 32;
    __shared__ uint32_t warp_cnt[NW][256];
    __shared__ uint32_t running[256];  // global destination of the next key of each digit
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    running[threadIdx.x] = offsets[(size_t)threadIdx.x * nblocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < NW; ++w) warp_cnt[w][threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RADIX_CHUNK;
    for (int r = 0; r < RADIX_ITEMS; ++r) {
        uint64_t i = base + (uint64_t)r * RADIX_THREADS + threadIdx.x;
        bool live = i < n;
        uint32_t key = live ? keys_in[i] : 0u;
        uint32_t val = live ? vals_in[i] : 0u;
        uint32_t digit = live ? ((key >> shift) & 255u) : 256u;  // 256: dead lanes match each other only
        unsigned peers = __match_any_sync(B200_FULL_MASK, digit);
        unsigned lt = peers & ((1u << lane) - 1u);
        uint32_t rank_in_warp = __popc(lt);
        if (live && lt == 0) warp_cnt[warp][digit] = __popc(peers);
        __syncthreads();
        if (live) {
            uint32_t off = running[digit] + rank_in_warp;
            for (int w = 0; w < warp; ++w) off += warp_cnt[w][digit];
            keys_out[off] = key;
            vals_out[off] = val;
        }
        __syncthreads();
        {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                tot += warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = 0;
            }
            running[threadIdx.x] += tot;
        }
        __syncthreads();
    }
}

// first k entries of the sorted ranking -> D/I (pads past the last valid row)
template <int METRIC>
__global__ void __launch_bounds__(256)
fullrank_emit_kernel(const uint32_t* __restrict__ inv_sorted, const uint32_t* __restrict__ rows_sorted,
                     uint64_t n, int64_t k, const int64_t* __restrict__ id_map, int64_t id_base,
                     float* __restrict__ D, int64_t* __restrict__ I) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < k; i += stride) {
        float dist = (METRIC == 0) ? -FLT_MAX : FLT_MAX;
        int64_t id = -1;
        if ((uint64_t)i < n) {
            uint32_t hi = ~inv_sorted[i];
            if (hi != 0u) {
                dist = b200_unord_f32(METRIC == 0 ? hi : ~hi);
                uint32_t row = rows_sorted[i];
                id = id_map ? id_map[row] : (int64_t)row + id_base;
            }
        }
        D[i] = dist;
        I[i] = id;
    }
}
