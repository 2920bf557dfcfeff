// rows.cuh — K1: row ingest at add time (optional L2 normalisation, optional fp32 -> bf16 store),
// and the counter-based synthetic row generator.
//
// K1 replaces memo's normalize() (memo_cli.py:131-135): n = sqrt(sum v^2) in fp32; n <= 1e-8 ->
// zero row; else v / n with a true (correctly rounded) division.  One warp per row; lane l owns
// the 4-element chunks l, l+32, ... and sums squares in ascending element order with fmaf, then a
// 16/8/4/2/1 xor butterfly (oracle/flat_oracle.c: oracle_normalize_device_order).
#pragma once
#include <climits>
#include "common.cuh"

struct IngestParams {
    const float* src;      // [n, d] fp32, dense
    uint8_t* dst;          // row storage, first destination row
    uint64_t pitch_bytes;  // destination pitch
    uint64_t n;
    int d;
    int d_pad;             // destination elements per row (zero padded)
};

template <int STORE>
__device__ __forceinline__ void ingest_store4(uint8_t* dst, int c, const float (&v)[4]) {
    if (STORE == 0) {
        *reinterpret_cast<float4*>(dst + 16 * (size_t)c) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
        __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
        uint2 w;
        w.x = *reinterpret_cast<uint32_t*>(&a);
        w.y = *reinterpret_cast<uint32_t*>(&b);
        *reinterpret_cast<uint2*>(dst + 8 * (size_t)c) = w;
    }
}

// REGS > 0: the row (d % 4 == 0, d <= 128 * REGS, 16-byte aligned source and destination) is loaded
// ONCE into REGS float4 registers per lane, the norm is reduced, and the scaled row is stored from the
// registers — one HBM read and one write per element, all of a lane's loads in flight together.
// REGS == 0: general two-pass form (odd d, unaligned pointers, very long rows).
template <int STORE, int NORMALIZE, int VEC, int REGS>
__global__ void __launch_bounds__(256) ingest_rows_kernel(const IngestParams p) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps_total = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const int nchunk = (p.d + 3) >> 2;
    for (uint64_t row = warp_global; row < p.n; row += warps_total) {
        const float* src = p.src + row * (uint64_t)p.d;
        uint8_t* dst = p.dst + row * p.pitch_bytes;
        if (REGS > 0) {
            float4 v[REGS > 0 ? REGS : 1];
#pragma unroll
            for (int j = 0; j < REGS; ++j) {
                const int c = lane + 32 * j;
                v[j] = c < nchunk ? *reinterpret_cast<const float4*>(src + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float nrm = 1.0f;
            bool zero = false;
            if (NORMALIZE) {
                float acc = 0.0f;
#pragma unroll
                for (int j = 0; j < REGS; ++j) {  // same order as the two-pass form: chunks lane, lane+32, ...
                    if (lane + 32 * j < nchunk) {
                        acc = fmaf(v[j].x, v[j].x, acc);
                        acc = fmaf(v[j].y, v[j].y, acc);
                        acc = fmaf(v[j].z, v[j].z, acc);
                        acc = fmaf(v[j].w, v[j].w, acc);
                    }
                }
                acc = warp_sum_xor(acc);
                nrm = __fsqrt_rn(acc);
                zero = ((double)nrm <= 1e-8);
            }
#pragma unroll
            for (int j = 0; j < REGS; ++j) {
                const int c = lane + 32 * j;
                if (c < (p.d_pad >> 2)) {
                    float o[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                    if (NORMALIZE) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = zero ? 0.0f : __fdiv_rn(o[e], nrm);
                    }
                    ingest_store4<STORE>(dst, c, o);
                }
            }
            continue;
        }
        float scale_den = 1.0f;
        bool zero = false;
        if (NORMALIZE) {
            float acc = 0.0f;
            for (int c = lane; c < nchunk; c += 32) {
                if (VEC) {
                    float4 v = *reinterpret_cast<const float4*>(src + 4 * c);
                    acc = fmaf(v.x, v.x, acc);
                    acc = fmaf(v.y, v.y, acc);
                    acc = fmaf(v.z, v.z, acc);
                    acc = fmaf(v.w, v.w, acc);
                } else {
                    for (int e = 4 * c; e < 4 * c + 4 && e < p.d; ++e) {
                        float v = src[e];
                        acc = fmaf(v, v, acc);
                    }
                }
            }
            acc = warp_sum_xor(acc);
            float nrm = __fsqrt_rn(acc);
            zero = ((double)nrm <= 1e-8);  // memo_cli.py:133 compares against the double 1e-8
            scale_den = nrm;
        }
        for (int c = lane; c < ((p.d_pad + 3) >> 2); c += 32) {
            float v[4];
            if (VEC && 4 * c + 3 < p.d) {
                float4 t = *reinterpret_cast<const float4*>(src + 4 * c);
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (4 * c + j < p.d) ? src[4 * c + j] : 0.0f;
            }
            if (NORMALIZE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = zero ? 0.0f : __fdiv_rn(v[j], scale_den);
            }
            if (STORE == 0 && (p.pitch_bytes & 15) != 0) {  // dense destination with d % 4 != 0 (normalised query scratch)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (4 * c + j < p.d) reinterpret_cast<float*>(dst)[4 * c + j] = v[j];
            } else {
                ingest_store4<STORE>(dst, c, v);
            }
        }
    }
}

// u(seed,row,col) in [-1,1) written as dense fp32 [n,d]
__global__ void __launch_bounds__(256) synth_rows_kernel(float* out, uint64_t n, uint32_t d,
                                                         uint64_t seed, uint64_t first_row) {
    const uint64_t total = n * (uint64_t)d;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        uint64_t row = i / d;
        uint32_t col = (uint32_t)(i - row * d);
        out[i] = b200_synth_value(seed, first_row + row, d, col);
    }
}

__global__ void __launch_bounds__(256) iota_ids_kernel(int64_t* out, uint64_t n, int64_t first) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = first + (int64_t)i;
}


// ---------------------------------------------------------------------------------------------
// id filter -> row bitmap (filtered search by record id).  The allowed ids are
// scattered into a bitmap over the ID range [base, base + range) (one atomicOr each), then every
// row looks its own id up in it and one ballot per 32 rows writes the row bitmap word the scan
// kernels test.  Sparse id spaces use the sorted-list form (binary search per row).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ids_minmax_kernel(const int64_t* __restrict__ ids, uint64_t n, long long* out) {
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        long long v = ids[i];
        lo = v < lo ? v : lo;
        hi = v > hi ? v : hi;
    }
    for (int o = 16; o; o >>= 1) {
        long long a = __shfl_xor_sync(0xffffffffu, lo, o), b = __shfl_xor_sync(0xffffffffu, hi, o);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

__global__ void __launch_bounds__(256) scatter_allowed_kernel(const int64_t* __restrict__ allowed, uint64_t m, int64_t base,
                                                              uint64_t range, uint32_t* __restrict__ bits) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const int64_t id = allowed[i];
        if (id < base) continue;
        const uint64_t off = (uint64_t)id - (uint64_t)base;
        if (off < range) atomicOr(&bits[off >> 5], 1u << (off & 31));
    }
}

// one thread per row, one ballot per warp = one row-bitmap word
__global__ void __launch_bounds__(256) gather_row_mask_kernel(const int64_t* __restrict__ ids, uint64_t n, int64_t base,
                                                              uint64_t range, const uint32_t* __restrict__ bits,
                                                              uint32_t* __restrict__ row_mask) {
    const uint64_t words = (n + 31) / 32;
    const uint64_t wstride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words; w += wstride) {
        const uint64_t r = w * 32 + lane;
        bool ok = false;
        if (r < n) {
            const int64_t id = ids[r];
            if (id >= base) {
                const uint64_t off = (uint64_t)id - (uint64_t)base;
                ok = off < range && ((bits[off >> 5] >> (off & 31)) & 1u);
            }
        }
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) row_mask[w] = word;
    }
}

__global__ void __launch_bounds__(256) gather_row_mask_sorted_kernel(const int64_t* __restrict__ ids, uint64_t n,
                                                                     const int64_t* __restrict__ sorted, uint64_t m,
                                                                     uint32_t* __restrict__ row_mask) {
    const uint64_t words = (n + 31) / 32;
    const uint64_t wstride = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    for (uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < words; w += wstride) {
        const uint64_t r = w * 32 + lane;
        bool ok = false;
        if (r < n) {
            const int64_t id = ids[r];
            uint64_t lo = 0, hi = m;
            while (lo < hi) {
                const uint64_t mid = (lo + hi) >> 1;
                if (sorted[mid] < id) lo = mid + 1;
                else hi = mid;
            }
            ok = lo < m && sorted[lo] == id;
        }
        const uint32_t word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) row_mask[w] = word;
    }
}
