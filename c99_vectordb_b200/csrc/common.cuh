// common.cuh — shared device helpers: result keys, PTX wrappers (mbarrier, cp.async.bulk),
// the counter-based synthetic generator.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <float.h>

#define B200_FULL_MASK 0xffffffffu

// ---------------------------------------------------------------------------------------------
// Result keys.  A candidate (score, row) is packed into one 64-bit key so that
//     key_a > key_b   <=>   a is better than b under the stated tie rule
// (score best-first, then smaller row first).  hi = monotone map of the fp32 score
// (inverted for L2 where smaller is better), lo = 0xFFFFFFFF - row.  key 0 is "empty":
// no valid score maps to hi == 0 (that would need the NaN bit pattern 0xFFFFFFFF).
// Valid candidates are s > -FLT_MAX (IP) / s < +FLT_MAX (L2): the faiss heaps start at
// -FLT_MAX / +FLT_MAX and replace on strict compare only, and NaN fails every compare
// [upstream faiss utils/Heap.h, ResultHandler.h — see DESIGN.md §4].
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t b200_ord_f32(float s) {
#ifdef __CUDA_ARCH__
    uint32_t u = (s == 0.0f) ? 0u : __float_as_uint(s);  // -0 and +0 are the same score
#else
    union { float f; uint32_t u; } c; c.f = s;
    uint32_t u = (s == 0.0f) ? 0u : c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float b200_unord_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
template <int METRIC>
__host__ __device__ __forceinline__ bool b200_score_valid(float s) {
    return METRIC == 0 ? (s > -FLT_MAX) : (s < FLT_MAX);
}
template <int METRIC>
__host__ __device__ __forceinline__ uint32_t b200_key_hi(float s) {
    uint32_t o = b200_ord_f32(s);
    return METRIC == 0 ? o : ~o;
}
template <int METRIC>
__host__ __device__ __forceinline__ uint64_t b200_make_key(float s, uint32_t row) {
    return ((uint64_t)b200_key_hi<METRIC>(s) << 32) | (uint64_t)(0xFFFFFFFFu - row);
}
__host__ __device__ __forceinline__ float b200_key_score(uint64_t key, int metric) {
    uint32_t hi = (uint32_t)(key >> 32);
    return b200_unord_f32(metric == 0 ? hi : ~hi);
}
__host__ __device__ __forceinline__ uint32_t b200_key_row(uint64_t key) {
    return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
}

// ---------------------------------------------------------------------------------------------
// Counter-based synthetic generator.  u(seed,row,col) in [-1,1), exact in fp32:
// a 24-bit integer times 2^-23 minus 1.  Identical integer arithmetic in oracle/flat_oracle.c.
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t b200_synth_bits(uint64_t seed, uint64_t ctr) {
    uint64_t z = ctr + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
__host__ __device__ __forceinline__ float b200_synth_value(uint64_t seed, uint64_t row, uint32_t d,
                                                            uint32_t col) {
    uint32_t b = b200_synth_bits(seed, row * (uint64_t)d + col);
    return (float)(b >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP), completion on an mbarrier.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes,
                                              uint32_t bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// streaming 128-bit global load, no L1 allocation
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// Programmatic dependent launch (no-ops in a kernel launched without the attribute).
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float warp_sum_xor(float v) {
    v += __shfl_xor_sync(B200_FULL_MASK, v, 16);
    v += __shfl_xor_sync(B200_FULL_MASK, v, 8);
    v += __shfl_xor_sync(B200_FULL_MASK, v, 4);
    v += __shfl_xor_sync(B200_FULL_MASK, v, 2);
    v += __shfl_xor_sync(B200_FULL_MASK, v, 1);
    return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        uint64_t o = __shfl_xor_sync(B200_FULL_MASK, v, m);
        v = o < v ? o : v;
    }
    return v;
}
#endif  // __CUDACC__
