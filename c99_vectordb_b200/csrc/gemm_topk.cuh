// gemm_topk.cuh — K3: batched queries as a dense contraction on the 5th-gen tensor cores.
//
// Replaces index.search(x[nq,d], k) for large nq (memo_cli.py:292; faiss's BLAS path for nq >= 20
// [upstream]).  10k queries x 10M rows x 768 is 1.5e14 flop: tensor-core work, not a GEMV.
//
// tcgen05 has no fp32 input kind, so one tensor-core pass cannot give fp32-exact ids.  The path is
// therefore "approximate, then prove":
//   1. shadow copies: rows and queries rounded to bf16 (K-major), row/query norms kept in fp32.
//   2. THIS kernel: S' = Q' * DB'^T tile by tile (UMMA 128 x 256 x 16, accumulators in TMEM,
//      operands staged by TMA with 128-byte swizzle through a 4-stage mbarrier ring), with a fused
//      epilogue that never writes scores to HBM:
//        MODE_TILEMAX : max score per (query, 32-row group) over a strided sample of tiles — used
//                       to pick each query's emission threshold theta;
//        MODE_EMIT    : rows whose approximate score exceeds theta[query] are appended to the
//                       query's candidate list (expected ~1e-4 of all scores).
//   3. rerank kernel: exact fp32 scores of the candidates with the SAME arithmetic as the scan
//      kernel (bit-identical to the nq < 20 path), top-k under the tie rule, and a certificate:
//      every row not emitted has exact score <= theta + eps (eps = proven bf16 rounding bound), so
//      the result is exact iff the k-th best exact score beats theta + eps.  Uncertified queries
//      are re-run through the exact scan kernel.
//
// Warp roles (6 warps): 0 = TMA producer, 1 = MMA issuer (+ TMEM allocator), 2..5 = epilogue
// (warp w reads TMEM lanes 32*(w%4)..+31: one query per thread).  Producer and issuer run their loops warp-uniformly
// and predicate the TMA / tcgen05 instructions on one elected lane (see elect_one_pred).  CG = 2 pairs two CTAs (SMs) on one
// 256 x 256 tile with cta_group::2; see gemm_topk_kernel.
#pragma once
#include <cuda.h>

#include "common.cuh"

#define G3_BLOCK_M 128
#define G3_BLOCK_N 256
#define G3_BLOCK_K 64   // bf16 elements = 128 bytes = one swizzle atom row
#define G3_UMMA_K 16
#define G3_THREADS 192
#define G3_A_BYTES (G3_BLOCK_M * G3_BLOCK_K * 2)
#define G3_MODE_TILEMAX 0
#define G3_MODE_EMIT 1

struct GemmParams {
    uint32_t nq;            // real queries
    uint32_t m_tiles;       // ceil(nq / 128)
    uint64_t n;             // database rows
    uint32_t k_blocks;      // d_pad64 / 64
    uint32_t tile_first;    // first 256-row tile of this pass (database row = tile * 256)
    uint32_t tile_stride;   // tile step (sampling)
    uint32_t tile_count;    // tiles in this pass
    uint32_t src_tile_first;   // where pass tile t sits in the bf16 matrix behind tm_db: src_tile_first + t * src_tile_stride
    uint32_t src_tile_stride;  // (the resident shadow: same as tile_first / tile_stride; a streamed chunk: 0 / 1)
    uint32_t chunk_tiles;   // tiles per work unit (L2 reuse window)
    int mode;
    const float* theta;     // [nq] emission thresholds (EMIT)
    unsigned int* cand_count;  // [nq]
    uint32_t* cand_rows;    // [nq, cand_cap]
    uint32_t cand_cap;
    float* tilemax;         // [nq, tile_count * 8]: one maximum per 32-row group (TILEMAX)
    const uint32_t* row_mask;  // optional row bitmap (filtered search): excluded rows score -inf
};

// ---- PTX: TMA tensor load, tcgen05 ------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld_32x32b_x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, tied to the registers of the load it completes: nothing that reads v[] can be scheduled above
// it (a second tcgen05.ld may stay in flight across this point — the epilogue's software pipeline)
__device__ __forceinline__ void tc_wait_ld_for(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :
                 : "memory");
}
// maximum of 32 accumulator columns as a balanced tree (the compiler pairs the levels into 3-input FMNMX3)
__device__ __forceinline__ float g3_max32(const uint32_t (&v)[32]) {
    float m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
        m[i] = fmaxf(fmaxf(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1])),
                     fmaxf(__uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3])));
    return fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7])));
}
// the same with the columns whose bit in `ok` is clear left out (ragged last tile, filtered search)
__device__ __forceinline__ float g3_max32_where(const uint32_t (&v)[32], uint32_t ok) {
    float m = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) m = fmaxf(m, ((ok >> j) & 1u) ? __uint_as_float(v[j]) : -INFINITY);
    return m;
}

// K-major operand tile in shared memory written by TMA with CU_TENSOR_MAP_SWIZZLE_128B:
// rows of 128 bytes, 8-row groups 1024 bytes apart.  Descriptor fields (cute::UMMA::SmemDescriptor):
// start>>4 [0,14), LBO>>4 [16,30) (unused for swizzled K-major, set to 1), SBO>>4 [32,46) = 1024>>4,
// version=1 [46,48), layout SWIZZLE_128B = 2 [61,64).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16 instruction descriptor: D=F32 (1<<4), A=BF16 (1<<7), B=BF16 (1<<10), both K-major,
// N>>3 at [17,23), M>>4 at [24,29).
__device__ __forceinline__ uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// ---- 2-CTA helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_alloc_cg(uint32_t smem_dst, uint32_t ncols) {
    if (CG == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
}
template <int CG>
__device__ __forceinline__ void tc_dealloc_cg(uint32_t taddr, uint32_t ncols) {
    if (CG == 1)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
    else
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// Warp-uniform issue.  Inside a divergent `if (lane == 0)` region ptxas has to move every operand of a tcgen05
// instruction into uniform registers through an ELECT / R2UR waterfall (about 14 instructions and a loop per MMA —
// measurable once an MMA takes 32 cycles).  Here every lane of the MMA warp runs the loop, the operands are computed
// in the uniform datapath, and the instructions are predicated on one elected lane.
__device__ __forceinline__ uint32_t elect_one_pred() {
    uint32_t e;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(e));
    return e;
}
// elected-lane forms of the TMA producer's instructions (the producer warp runs its loop warp-uniformly too)
__device__ __forceinline__ void mbar_arrive_expect_tx_if(uint32_t elected, uint32_t bar, uint32_t bytes) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "setp.ne.b32 e, %2, 0;\n\t"
        "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}"
        ::"r"(bar), "r"(bytes), "r"(elected)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_if(uint32_t elected, uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
        ::"r"(dst_smem), "l"(map), "r"(c0), "r"(c1), "r"(bar), "r"(elected)
        : "memory");
}
// TMA load issued by either CTA of a pair; completion bytes are credited to the mbarrier of the pair's leader
// (mbar is a shared::cluster address inside the leader CTA)
__device__ __forceinline__ void tma_load_2d_cg2_if(uint32_t elected, uint32_t dst_smem, const CUtensorMap* map, int c0, int c1, uint32_t mbar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n\t}"
        ::"r"(dst_smem), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(elected)
        : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_mma_f16_cg_if(uint32_t elected, uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred p, e;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 e, %5, 0;\n\t"
            "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p, e;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "setp.ne.b32 e, %5, 0;\n\t"
            "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
            : "memory");
}
template <int CG>
__device__ __forceinline__ void tc_commit_cg_if(uint32_t elected, uint32_t bar) {
    if (CG == 1)
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "setp.ne.b32 e, %1, 0;\n\t"
            "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
            ::"r"(bar), "r"(elected)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred e;\n\t"
            "setp.ne.b32 e, %2, 0;\n\t"
            "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
            ::"r"(bar), "h"((uint16_t)3), "r"(elected)
            : "memory");
}

template <int CG>
struct G3Cfg {
    static constexpr uint32_t kBRows = G3_BLOCK_N / CG;            // rows of the N tile staged by one CTA
    static constexpr uint32_t kBBytes = kBRows * G3_BLOCK_K * 2;
    static constexpr uint32_t kStageBytes = G3_A_BYTES + kBBytes;  // 48 KB (CG=1) / 32 KB (CG=2)
    static constexpr uint32_t kStages = CG == 1 ? 4 : 6;
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + 1024 + 256;
};

// CG = 1: one CTA per 128 x 256 tile.  CG = 2: a CTA pair (cluster of 2, cta_group::2) per
// 256 x 256 tile — each CTA stages its own 128 queries and HALF of the 256 rows, the pair's tensor
// cores share the row operand, which halves shared-memory traffic per SM (the cta_group::1 form is
// capped near 2/3 of peak by the 128 B/cycle smem port: 96 B/cycle of operand reads + 96 B/cycle of
// TMA fills).  Only the pair's leader (rank 0) issues MMAs; both CTAs run TMA and the epilogue.
template <int CG, bool MASKED, int MODE>
__global__ void __launch_bounds__(G3_THREADS, 1)
gemm_topk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                 const GemmParams p) {
    using Cfg = G3Cfg<CG>;
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment for the swizzle atoms (identical offset in both CTAs of a pair)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * Cfg::kStages + 4);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (Cfg::kStages + s); };
    auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * Cfg::kStages + b); };
    auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * Cfg::kStages + 2 + b); };

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 1 ? 0u : cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t group_id = blockIdx.x / CG, groups = gridDim.x / CG;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_db);
        for (uint32_t s = 0; s < Cfg::kStages; ++s) {
            mbar_init(full_bar(s), 1);   // the leader's arrive.expect_tx (bytes of both CTAs land here)
            mbar_init(empty_bar(s), 1);  // one (multicast) tcgen05.commit
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 4 * CG);  // one arrive per epilogue warp of every CTA in the pair
        }
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc_cg<CG>(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything targets them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t chunks = (p.tile_count + p.chunk_tiles - 1) / p.chunk_tiles;
    const uint32_t m_groups = (p.m_tiles + CG - 1) / CG;  // 128*CG queries per work unit
    const uint32_t units = chunks * m_groups;  // unit u -> (chunk = u / m_groups, m group = u % m_groups):
                                               // CTAs running together share a chunk of rows in L2

    if (warp == 0) {
        // ===== TMA producer (every CTA) =====
        {
            const uint32_t elected = elect_one_pred();
            uint32_t stage = 0, phase = 0;
            // CG == 2: completion bytes are credited to the LEADER's full barrier
            for (uint32_t u = group_id; u < units; u += groups) {
                const uint32_t chunk = u / m_groups, mg = u - chunk * m_groups;
                const uint32_t mt = mg * CG + rank;
                const uint32_t t0 = chunk * p.chunk_tiles;
                const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
                for (uint32_t t = t0; t < t1; ++t) {
                    const uint32_t ntile = p.src_tile_first + t * p.src_tile_stride;
                    for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t a_dst = smem_base + stage * Cfg::kStageBytes;
                        const int kc = (int)(kb * G3_BLOCK_K);
                        if (CG == 1) {
                            mbar_arrive_expect_tx_if(elected, full_bar(stage), Cfg::kStageBytes);
                            tma_load_2d_if(elected, a_dst, &tm_q, kc, (int)(mt * G3_BLOCK_M), full_bar(stage));
                            tma_load_2d_if(elected, a_dst + G3_A_BYTES, &tm_db, kc, (int)(ntile * G3_BLOCK_N), full_bar(stage));
                        } else {
                            if (leader) mbar_arrive_expect_tx_if(elected, full_bar(stage), Cfg::kStageBytes * 2);
                            const uint32_t lead_bar = mapa_shared(full_bar(stage), 0);
                            tma_load_2d_cg2_if(elected, a_dst, &tm_q, kc, (int)(mt * G3_BLOCK_M), lead_bar);
                            tma_load_2d_cg2_if(elected, a_dst + G3_A_BYTES, &tm_db, kc, (int)(ntile * G3_BLOCK_N + rank * Cfg::kBRows), lead_bar);
                        }
                        if (++stage == Cfg::kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the pair's leader only) =====
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(G3_BLOCK_M * CG, G3_BLOCK_N);
            const uint32_t elected = elect_one_pred();
            uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0;
            for (uint32_t u = group_id; u < units; u += groups) {
                const uint32_t chunk = u / m_groups;
                const uint32_t t0 = chunk * p.chunk_tiles;
                const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
                for (uint32_t t = t0; t < t1; ++t) {
                    mbar_wait(tempty_bar(abuf), aphase ^ 1u);  // every epilogue warp has drained this accumulator
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + abuf * G3_BLOCK_N;
                    for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t a_addr = smem_base + stage * Cfg::kStageBytes;
                        const uint64_t adesc = umma_desc_k_sw128(a_addr);
                        const uint64_t bdesc = umma_desc_k_sw128(a_addr + G3_A_BYTES);
#pragma unroll
                        for (uint32_t k = 0; k < G3_BLOCK_K / G3_UMMA_K; ++k) {
                            // advance 32 bytes (16 bf16) inside the swizzle atom: +2 in the >>4 encoded address
                            tc_mma_f16_cg_if<CG>(elected, d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                        tc_commit_cg_if<CG>(elected, empty_bar(stage));  // smem slot free (in both CTAs) once these MMAs retire
                        if (kb + 1 == p.k_blocks) tc_commit_cg_if<CG>(elected, tfull_bar(abuf));  // accumulator complete
                        if (++stage == Cfg::kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    if (++abuf == 2) {
                        abuf = 0;
                        aphase ^= 1u;
                    }
                }
            }
        }
    } else {
        // ===== epilogue: one query per thread (every CTA reads its own 128 TMEM lanes) =====
        const uint32_t quarter = (uint32_t)warp & 3u;  // TMEM lanes this warp may read
        const uint32_t tempty_lead0 = CG == 1 ? tempty_bar(0) : mapa_shared(tempty_bar(0), 0);
        uint32_t abuf = 0, aphase = 0;
        for (uint32_t u = group_id; u < units; u += groups) {
            const uint32_t chunk = u / m_groups, mg = u - chunk * m_groups;
            const uint32_t mt = mg * CG + rank;
            const uint32_t t0 = chunk * p.chunk_tiles;
            const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
            const uint32_t qidx = mt * G3_BLOCK_M + quarter * 32u + (uint32_t)lane;
            const bool qlive = qidx < p.nq;
            float theta = INFINITY;
            if (MODE == G3_MODE_EMIT && qlive) theta = p.theta[qidx];
            for (uint32_t t = t0; t < t1; ++t) {
                const uint32_t ntile = p.tile_first + t * p.tile_stride;
                const uint64_t row0 = (uint64_t)ntile * G3_BLOCK_N;
                const uint32_t live_cols = (uint32_t)(p.n - row0 < G3_BLOCK_N ? p.n - row0 : G3_BLOCK_N);
                mbar_wait(tfull_bar(abuf), aphase);
                tc_fence_after();
                const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + abuf * G3_BLOCK_N;
                float gmax[G3_BLOCK_N / 32];
                // one 32-column group: its maximum, and (EMIT) the rare append of the rows above theta.
                // `ok` (live columns of a ragged last tile & the filter's mask word) is the same for every
                // thread of the warp, so the all-ones test is a uniform branch and the common case is 16
                // FMNMX3 + one compare per 32 scores.
                auto group = [&](const uint32_t (&v)[32], uint32_t g) {
                    const uint32_t c0 = g * 32u;
                    uint32_t ok = c0 >= live_cols ? 0u : (live_cols - c0 >= 32u ? 0xFFFFFFFFu : (1u << (live_cols - c0)) - 1u);
                    if (MASKED && ok) ok &= __ldg(p.row_mask + ((row0 + c0) >> 5));  // tiles start at multiples of 256 rows
                    float m;
                    if (ok == 0xFFFFFFFFu) m = g3_max32(v);
                    else if (ok == 0u) m = -INFINITY;  // rows past the end of the database (zero filled by TMA)
                    else m = g3_max32_where(v, ok);
                    if (MODE == G3_MODE_TILEMAX) gmax[g] = m;
                    if (MODE == G3_MODE_EMIT && m > theta) {
                        uint32_t hit = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) hit |= (__uint_as_float(v[j]) > theta) ? (1u << j) : 0u;
                        hit &= ok;
                        while (hit) {
                            const int j = __ffs(hit) - 1;
                            hit &= hit - 1;
                            unsigned pos = atomicAdd(p.cand_count + qidx, 1u);
                            if (pos < p.cand_cap) p.cand_rows[(size_t)qidx * p.cand_cap + pos] = (uint32_t)(row0 + c0 + j);
                        }
                    }
                };
                // software pipeline over the 8 column groups: the tcgen05.ld of group g+1 is in flight while
                // group g is reduced; the accumulator is handed back to the MMA warp as soon as the last
                // load has landed, before the last group is processed
                uint32_t va[32], vb[32];
                tc_ld_32x32b_x32(taddr0, va);
#pragma unroll
                for (uint32_t g = 0; g < G3_BLOCK_N / 32; g += 2) {
                    tc_wait_ld_for(va);
                    tc_ld_32x32b_x32(taddr0 + (g + 1) * 32u, vb);
                    group(va, g);
                    tc_wait_ld_for(vb);
                    if (g + 2 < G3_BLOCK_N / 32) {
                        tc_ld_32x32b_x32(taddr0 + (g + 2) * 32u, va);
                    } else {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (CG == 1) mbar_arrive(tempty_bar(abuf));
                            else mbar_arrive_cluster(tempty_lead0 + 8u * abuf);  // the leader's MMA warp waits for both CTAs
                        }
                    }
                    group(vb, g + 1);
                }
                if (MODE == G3_MODE_TILEMAX && qlive) {
                    float4* dst = reinterpret_cast<float4*>(p.tilemax + ((size_t)qidx * p.tile_count + t) * 8);
                    dst[0] = make_float4(gmax[0], gmax[1], gmax[2], gmax[3]);
                    dst[1] = make_float4(gmax[4], gmax[5], gmax[6], gmax[7]);
                }
                if (++abuf == 2) {
                    abuf = 0;
                    aphase ^= 1u;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();  // nobody leaves while its pair still touches its smem / TMEM
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc_cg<CG>(tmem_base, 512);
    }
}

// ---- K3 for small batches: rows as the M operand, queries resident ---------------------------------------------
// With at most a few hundred queries the 256 x 256 kernel above is bound by operand traffic, not by the tensor pipe:
// every 256-row tile re-fetches the whole query tile from L2 (as many bytes as the rows themselves), so a sweep of
// the 15 GB shadow moves 30 GB from L2 into shared memory and takes 3.4-4.1 ms where HBM would allow 2.1.  This
// form swaps the roles: the database rows are the M operand (256 rows per CTA pair, streamed by TMA through a
// ring), the queries are the N operand (N = nq rounded up to 16, N <= 256 while N/2 x kpad bf16 fit next to the
// ring) and are loaded into shared memory ONCE per CTA — the only stream is the shadow itself.  One
// tcgen05.mma.cta_group::2 (M = 256, N, K = 16) per 32 bytes of K; accumulators [row lane, query column] in TMEM,
// double buffered.  Epilogue: a thread owns one database row and compares its N scores with the queries' thresholds
// (32 at a time from shared memory); the per-(query, 32-row group) maxima of the threshold pass are warp
// reductions (redux.sync on the ordered-integer image).  Same candidate lists, same re-rank, same certificate.
#define G3T_THREADS 192
#define G3T_A_BYTES (128 * G3_BLOCK_K * 2)  // one CTA's half of a 256-row tile, one K block: 16 KB

struct GemmRowsParams {
    uint32_t nq;            // real queries
    uint32_t n_cols;        // N: queries padded to a multiple of 16 (<= 256)
    uint64_t n;             // database rows
    uint32_t k_blocks;      // kpad / 64
    uint32_t tile_first, tile_stride, tile_count;          // database tiles of this pass (as GemmParams)
    uint32_t src_tile_first, src_tile_stride;              // where they sit behind tm_db
    uint32_t stages;        // ring stages (host: what fits next to the resident queries)
    int mode;
    const float* theta;
    unsigned int* cand_count;
    uint32_t* cand_rows;
    uint32_t cand_cap;
    float* tilemax;         // [nq, tile_count * 8]
    const uint32_t* row_mask;
};

template <int MODE, bool MASKED>
__global__ void __launch_bounds__(G3T_THREADS, 1)
gemm_rows_topk_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_db,
                      const GemmRowsParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t half_cols = p.n_cols / 2;                      // query rows this CTA holds
    const uint32_t qblock_bytes = half_cols * G3_BLOCK_K * 2;     // one K block of them (multiple of 1024)
    const uint32_t q_bytes = p.k_blocks * qblock_bytes;
    const uint32_t ring_off = q_bytes;
    const uint32_t bar_off = ring_off + p.stages * G3T_A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + bar_off);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty, [2S+4]: queries resident
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 5);
    float* s_theta = reinterpret_cast<float*>(bars + 2 * p.stages + 6);  // [n_cols rounded up to 32], +inf past nq
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    const uint32_t S = p.stages;
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (S + s); };
    auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * S + b); };
    auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * S + 2 + b); };
    const uint32_t qres_bar = bar_base + 8u * (2 * S + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t group_id = blockIdx.x / 2, groups = gridDim.x / 2;
    // two accumulator buffers of a power-of-two number of columns >= N and >= 32 (the epilogue reads 32 columns at a time)
    const uint32_t tmem_cols = p.n_cols <= 32 ? 64u : (p.n_cols <= 64 ? 128u : (p.n_cols <= 128 ? 256u : 512u));
    const uint32_t acc_stride = tmem_cols / 2;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_q);
        tma_prefetch_desc(&tm_db);
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);  // four epilogue warps of each CTA of the pair
        }
        mbar_init(qres_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc_cg<2>(smem_u32(tmem_slot), tmem_cols);
    if (MODE == G3_MODE_EMIT)
        for (uint32_t j = threadIdx.x; j < ((p.n_cols + 31u) & ~31u); j += blockDim.x) s_theta[j] = j < p.nq ? p.theta[j] : INFINITY;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        {
            const uint32_t elected = elect_one_pred();
            // the queries, once: K block kb of this CTA's half of the query rows -> smem [kb]
            if (leader) mbar_arrive_expect_tx_if(elected, qres_bar, q_bytes * 2);
            const uint32_t lead_q = mapa_shared(qres_bar, 0);
            for (uint32_t kb = 0; kb < p.k_blocks; ++kb)
                tma_load_2d_cg2_if(elected, smem_base + kb * qblock_bytes, &tm_q, (int)(kb * G3_BLOCK_K), (int)(rank * half_cols), lead_q);
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = group_id; t < p.tile_count; t += groups) {
                const uint32_t src_tile = p.src_tile_first + t * p.src_tile_stride;
                for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    if (leader) mbar_arrive_expect_tx_if(elected, full_bar(stage), G3T_A_BYTES * 2);
                    const uint32_t lead_bar = mapa_shared(full_bar(stage), 0);
                    tma_load_2d_cg2_if(elected, smem_base + ring_off + stage * G3T_A_BYTES, &tm_db, (int)(kb * G3_BLOCK_K),
                                       (int)(src_tile * G3_BLOCK_N + rank * 128u), lead_bar);
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA) =====
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(256, p.n_cols);
            const uint32_t elected = elect_one_pred();
            mbar_wait(qres_bar, 0);
            tc_fence_after();
            uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0;
            for (uint32_t t = group_id; t < p.tile_count; t += groups) {
                mbar_wait(tempty_bar(abuf), aphase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + abuf * acc_stride;
                for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_k_sw128(smem_base + ring_off + stage * G3T_A_BYTES);
                    const uint64_t bdesc = umma_desc_k_sw128(smem_base + kb * qblock_bytes);
#pragma unroll
                    for (uint32_t k = 0; k < G3_BLOCK_K / G3_UMMA_K; ++k)
                        tc_mma_f16_cg_if<2>(elected, d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    tc_commit_cg_if<2>(elected, empty_bar(stage));
                    if (kb + 1 == p.k_blocks) tc_commit_cg_if<2>(elected, tfull_bar(abuf));
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (++abuf == 2) {
                    abuf = 0;
                    aphase ^= 1u;
                }
            }
        }
    } else {
        // ===== epilogue: one database row per thread =====
        const uint32_t quarter = (uint32_t)warp & 3u;
        const uint32_t tempty_lead0 = mapa_shared(tempty_bar(0), 0);
        const uint32_t groups32 = (p.n_cols + 31) / 32;
        uint32_t abuf = 0, aphase = 0;
        for (uint32_t t = group_id; t < p.tile_count; t += groups) {
            const uint32_t ntile = p.tile_first + t * p.tile_stride;
            const uint64_t row = (uint64_t)ntile * G3_BLOCK_N + rank * 128u + quarter * 32u + (uint32_t)lane;
            bool live = row < p.n;
            if (MASKED && live) live = (__ldg(p.row_mask + (row >> 5)) >> (row & 31)) & 1u;
            mbar_wait(tfull_bar(abuf), aphase);
            tc_fence_after();
            const uint32_t taddr0 = tmem_base + ((quarter * 32u) << 16) + abuf * acc_stride;
            for (uint32_t g = 0; g < groups32; ++g) {
                uint32_t v[32];
                tc_ld_32x32b_x32(taddr0 + g * 32u, v);
                tc_wait_ld_for(v);
                const uint32_t c0 = g * 32u;
                if (MODE == G3_MODE_TILEMAX) {
                    // per query column: the maximum over this warp's 32 rows = one 32-row group of the tile
                    float keep = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float sc = __uint_as_float(v[j]);
                        const uint32_t o = (live && sc == sc) ? b200_ord_f32(sc) : 0u;  // NaN scores never count
                        const uint32_t mx = __reduce_max_sync(B200_FULL_MASK, o);
                        if (lane == j) keep = mx ? b200_unord_f32(mx) : -INFINITY;
                    }
                    const uint32_t qidx = c0 + (uint32_t)lane;
                    if (qidx < p.nq) p.tilemax[((size_t)qidx * p.tile_count + t) * 8 + rank * 4u + quarter] = keep;
                } else {
                    if (live) {
                        uint32_t hit = 0;
#pragma unroll
                        for (int j4 = 0; j4 < 8; ++j4) {
                            const float4 th = *reinterpret_cast<const float4*>(s_theta + c0 + 4 * j4);
                            hit |= (__uint_as_float(v[4 * j4 + 0]) > th.x ? 1u : 0u) << (4 * j4 + 0);
                            hit |= (__uint_as_float(v[4 * j4 + 1]) > th.y ? 1u : 0u) << (4 * j4 + 1);
                            hit |= (__uint_as_float(v[4 * j4 + 2]) > th.z ? 1u : 0u) << (4 * j4 + 2);
                            hit |= (__uint_as_float(v[4 * j4 + 3]) > th.w ? 1u : 0u) << (4 * j4 + 3);
                        }
                        while (hit) {  // rare
                            const int j = __ffs(hit) - 1;
                            hit &= hit - 1;
                            const uint32_t qidx = c0 + (uint32_t)j;  // < nq: padded columns carry theta = +inf
                            const unsigned pos = atomicAdd(p.cand_count + qidx, 1u);
                            if (pos < p.cand_cap) p.cand_rows[(size_t)qidx * p.cand_cap + pos] = (uint32_t)row;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_lead0 + 8u * abuf);
            if (++abuf == 2) {
                abuf = 0;
                aphase ^= 1u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc_cg<2>(tmem_base, tmem_cols);
    }
}

// ---- shadow copies --------------------------------------------------------------------------------
// fp32 rows (pitch bytes) -> bf16 K-major rows of kpad elements (zero padded) + fp32 squared norms
// + a running maximum of the row norm (for the certificate's error bound).
// aug: 0 = none; 1 = database row for L2: columns d, d+1 hold h = |y|^2 / 2 split into two bf16
// (hi + lo, residual <= 2^-16 h); 2 = query for L2: columns d, d+1 hold -1.  With these the same
// GEMM yields q.y - |y|^2/2 = (|q|^2 - L2) / 2: larger is better, exactly like IP.
// gather_stride > 1: output row r is database row ((r >> 8) * gather_stride << 8) + (r & 255) — the sampled 256-row
// tiles of the threshold pass packed back to back (streamed shadow).  fp32 rows with d % 4 == 0 take the vector
// path: 128-bit loads marked evict-first in L2 (the rows are read once; the bf16 chunk written here is what the
// GEMM re-reads), 64-bit bf16x4 stores.
__global__ void __launch_bounds__(256)
shadow_rows_kernel(const uint8_t* __restrict__ rows, uint64_t pitch, int store, uint64_t n_out, uint64_t n_src, uint32_t gather_stride,
                   int d, int kpad, int aug, __nv_bfloat16* __restrict__ out, float* __restrict__ norm2,
                   unsigned int* __restrict__ max_norm2_bits) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t warps_total = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec = store == 0 && (d & 3) == 0 && (pitch & 15) == 0;
    const uint64_t pol = l2_policy_evict_first();
    float wmax = 0.0f;
    for (uint64_t r = warp_global; r < n_out; r += warps_total) {
        const uint64_t sr = gather_stride > 1 ? ((((r >> 8) * gather_stride) << 8) + (r & 255)) : r;
        __nv_bfloat16* dst = out + r * (uint64_t)kpad;
        float acc = 0.0f;
        if (sr >= n_src) {  // past the end of the database: a zero row (never a candidate: the epilogue masks it)
            for (int c = lane; c < kpad; c += 32) dst[c] = __float2bfloat16_rn(0.0f);
            continue;
        }
        const uint8_t* src = rows + sr * pitch;
        if (vec) {
            // all of this lane's 128-bit loads of the row are issued before the first is used (8 per pass: rows of up
            // to 1024 columns in one pass), so a warp keeps up to 4 KB in flight
            for (int base4 = 0; base4 < (kpad >> 2); base4 += 256) {
                float4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c4 = base4 + lane + 32 * u;
                    v[u] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                    if (c4 < (kpad >> 2) && 4 * c4 < d)
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                                     : "l"(src + 16 * (size_t)c4), "l"(pol));
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c4 = base4 + lane + 32 * u;
                    if (c4 >= (kpad >> 2)) break;
                    acc = fmaf(v[u].x, v[u].x, acc);
                    acc = fmaf(v[u].y, v[u].y, acc);
                    acc = fmaf(v[u].z, v[u].z, acc);
                    acc = fmaf(v[u].w, v[u].w, acc);
                    __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(dst + 4 * c4) = pk;
                }
            }
        } else {
            for (int c = lane; c < kpad; c += 32) {
                float v = 0.0f;
                if (c < d) v = store == 0 ? reinterpret_cast<const float*>(src)[c] : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[c]);
                acc = fmaf(v, v, acc);
                dst[c] = __float2bfloat16_rn(v);
            }
        }
        acc = warp_sum_xor(acc);
        __syncwarp();
        if (lane == 0) {
            if (norm2) norm2[r] = acc;
            if (aug == 1) {
                const float h = 0.5f * acc;
                const __nv_bfloat16 hi = __float2bfloat16_rn(h);
                dst[d] = hi;
                dst[d + 1] = __float2bfloat16_rn(h - __bfloat162float(hi));
            } else if (aug == 2) {
                dst[d] = __float2bfloat16_rn(-1.0f);
                dst[d + 1] = __float2bfloat16_rn(-1.0f);
            }
        }
        wmax = fmaxf(wmax, acc);
    }
    if (lane == 0 && max_norm2_bits) atomicMax(max_norm2_bits, __float_as_uint(wmax));  // non-negative floats order as uints
}

// ---- threshold selection: theta[q] ~ rank-th largest of tilemax[q, 0..T) ----------------------------
// Two-level radix select on the monotone integer image of the scores: a 256-bin histogram of the top
// byte locates the bin holding the rank-th largest value, a second histogram of the next byte inside
// that bin refines it; theta is the LOWER edge of the selected sub-bin, i.e. never above the true
// rank-th value (a slightly lower threshold only emits a few more candidates).  One CTA per query.
__global__ void __launch_bounds__(256)
select_theta_kernel(const float* __restrict__ tilemax, uint32_t T, uint32_t rank, float* __restrict__ theta) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t sel_bin, sel_rank;
    const float* src = tilemax + (size_t)blockIdx.x * T;
    uint32_t prefix = 0, want = rank < T ? rank : T - 1;  // 0-based rank among values sorted descending
    for (int level = 0; level < 2; ++level) {
        hist[threadIdx.x] = 0;
        __syncthreads();
        const int shift = 24 - 8 * level;
        for (uint32_t i = threadIdx.x; i < T; i += blockDim.x) {
            uint32_t o = b200_ord_f32(src[i]);
            if (level == 0 || (o >> 24) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t acc = 0;
            int b = 255;
            for (; b > 0; --b) {  // walk from the largest values down
                if (acc + hist[b] > want) break;
                acc += hist[b];
            }
            sel_bin = (uint32_t)b;
            sel_rank = want - acc;
        }
        __syncthreads();
        if (level == 0) prefix = sel_bin;
        else prefix = (prefix << 8) | sel_bin;
        want = sel_rank;
        __syncthreads();
    }
    if (threadIdx.x == 0) theta[blockIdx.x] = b200_unord_f32(prefix << 16);  // lower edge of the 16-bit prefix bucket
}

// ---- exact re-rank + certificate -------------------------------------------------------------------
// One CTA per query.  Every candidate row is re-scored in fp32 with exactly the scan kernel's
// arithmetic (lane l owns 16-byte chunks l, l+32, ...; fmaf in ascending element order; xor
// butterfly), so the distances are bit-identical to what the nq < 20 path returns.
struct RerankParams {
    const uint8_t* rows;
    uint64_t pitch_bytes;
    uint32_t nvec;
    int store;             // 0 fp32 rows, 1 bf16 rows
    int d;
    int qstride;           // floats per staged query (multiple of 8)
    int lpr;               // lanes sharing a row in the scan arithmetic (32, or 16 for short rows)
    const float* q;        // [nq, d] fp32 queries (the originals, not the bf16 shadow)
    const uint32_t* cand_rows;
    const unsigned int* cand_count;
    uint32_t cand_cap;
    const float* theta;    // [nq]
    const float* qnorm2;   // [nq]
    const unsigned int* max_norm2_bits;
    float eps_rel;         // proven relative bound on |approx - exact| / (|q| |y|)
    uint64_t n;            // database rows
    int k;
    const int64_t* id_map;
    float* D;
    int64_t* I;
    int* certified;        // [nq] 1 = result proven exact, 0 = must be recomputed by the exact scan
    float* bound;          // optional [nq] (row-sharded batches): no row outside the list scores better than this
};

template <int METRIC>
__global__ void __launch_bounds__(256) rerank_kernel(const RerankParams p) {
    extern __shared__ __align__(16) uint8_t sm[];
    const uint32_t qi = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* qs = reinterpret_cast<float*>(sm);
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm + (size_t)p.qstride * 4);
    const uint32_t count = p.cand_count[qi];
    const uint32_t c = count < p.cand_cap ? count : p.cand_cap;
    uint32_t m = 2;
    while (m < c) m <<= 1;
    for (int i = threadIdx.x; i < p.qstride; i += blockDim.x) qs[i] = i < p.d ? p.q[(size_t)qi * p.d + i] : 0.0f;
    for (uint32_t i = c + threadIdx.x; i < m; i += blockDim.x) keys[i] = 0ull;
    __syncthreads();
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    for (uint32_t j = warp; j < c; j += nw) {
        const uint32_t row = p.cand_rows[(size_t)qi * p.cand_cap + j];
        const uint8_t* rp = p.rows + (uint64_t)row * p.pitch_bytes;
        float acc = 0.0f;
        // lanes >= lpr stay at +0: adding them in the 32-lane butterfly below is exact, so the result
        // equals the lpr-lane butterfly of the scan kernel
        auto step = [&](uint32_t bits, float qv) {  // same element arithmetic as scan_topk.cuh acc1<>
            const float v = __uint_as_float(bits);
            if (METRIC == 0) {
                acc = fmaf(v, qv, acc);
            } else {
                const float t = v - qv;
                acc = fmaf(t, t, acc);
            }
        };
        auto chunk = [&](const uint4& raw, uint32_t ch) {
            if (p.store == 0) {
                float4 qv = q4[ch];
                step(raw.x, qv.x); step(raw.y, qv.y); step(raw.z, qv.z); step(raw.w, qv.w);
            } else {
                float4 qa = q4[2 * ch], qb = q4[2 * ch + 1];
                step(raw.x << 16, qa.x); step(raw.x & 0xffff0000u, qa.y);
                step(raw.y << 16, qa.z); step(raw.y & 0xffff0000u, qa.w);
                step(raw.z << 16, qb.x); step(raw.z & 0xffff0000u, qb.y);
                step(raw.w << 16, qb.z); step(raw.w & 0xffff0000u, qb.w);
            }
        };
        if (lane < p.lpr) {
            // up to eight 16-byte loads of the row in flight per lane before the first is used (a candidate row is a
            // random 3 KB read: latency, not arithmetic, bounds this kernel); the sums keep the scan's order
            for (uint32_t base = lane; base < p.nvec; base += 8 * p.lpr) {
                uint4 raw[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t ch = base + u * p.lpr;
                    if (ch < p.nvec) raw[u] = ldg_nc_v4(rp + (size_t)ch * 16);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const uint32_t ch = base + u * p.lpr;
                    if (ch < p.nvec) chunk(raw[u], ch);
                }
            }
        }
        acc = warp_sum_xor(acc);
        if (lane == 0) keys[j] = b200_score_valid<METRIC>(acc) ? b200_make_key<METRIC>(acc, row) : 0ull;
    }
    // descending bitonic sort of the candidate keys
    for (uint32_t size = 2; size <= m; size <<= 1)
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t x = keys[lo], y = keys[hi];
                if (desc ? (x < y) : (x > y)) {
                    keys[lo] = y;
                    keys[hi] = x;
                }
            }
        }
    __syncthreads();
    const int k = p.k;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = (uint32_t)i < m ? keys[i] : 0ull;
        if ((uint32_t)i >= c) key = 0ull;
        float dist = (METRIC == 0) ? -FLT_MAX : FLT_MAX;
        int64_t id = -1;
        if (key != 0ull) {
            dist = b200_key_score(key, METRIC);
            uint32_t row = b200_key_row(key);
            id = p.id_map ? p.id_map[row] : (int64_t)row;
        }
        p.D[(size_t)qi * k + i] = dist;
        p.I[(size_t)qi * k + i] = id;
    }
    if (threadIdx.x == 0) {
        // Certificate.  A row that was not emitted has approx <= theta, hence exact <= theta + eps.
        // The answer is proven iff the k-th best exact score is strictly above that (strict: an
        // equal score with a smaller row would win the tie).
        const uint64_t want = (uint64_t)k < p.n ? (uint64_t)k : p.n;
        bool ok = (count <= p.cand_cap) && (c >= want);
        if (ok && want > 0) {
            uint64_t kth = keys[want - 1];
            if (kth == 0ull) {
                ok = false;
            } else {
                const float maxn = sqrtf(__uint_as_float(*p.max_norm2_bits));
                const float qn2 = p.qnorm2[qi], qn = sqrtf(qn2);
                const float eps = p.eps_rel * qn * maxn;
                if (METRIC == 0) {
                    ok = b200_key_score(kth, METRIC) > p.theta[qi] + eps;
                } else {
                    // the GEMM ranks a = q.y - |y|^2/2 = (|q|^2 - L2)/2; a row that was not emitted has
                    // a <= theta + eps_a, i.e. L2 >= |q|^2 - 2 (theta + eps_a).  eps_a adds the hi/lo split
                    // residual of |y|^2/2; slack covers the fp32 rounding of |q|^2 and of both L2 sums.
                    const float eps_a = eps + 2e-5f * maxn * maxn;
                    const float slack = 4.0f * (float)p.d * 6e-8f * (qn + maxn) * (qn + maxn);
                    ok = b200_key_score(kth, METRIC) < qn2 - 2.0f * (p.theta[qi] + eps_a) - slack;
                }
            }
        }
        p.certified[qi] = ok ? 1 : 0;
        if (p.bound) {
            // Row-sharded batch: the certificate is taken after the merge, over all shards.  A row of this shard
            // that is not in the list was either not emitted — exact score no better than theta + eps (IP) /
            // |q|^2 - 2 (theta + eps_a) - slack (L2) — or cut off behind k better entries of this very list.
            // An overflowed candidate list proves nothing: the bound says so.
            const float maxn = sqrtf(__uint_as_float(*p.max_norm2_bits));
            const float qn2 = p.qnorm2[qi], qn = sqrtf(qn2);
            const float eps = p.eps_rel * qn * maxn;
            float b;
            if (METRIC == 0) {
                b = count <= p.cand_cap ? p.theta[qi] + eps : INFINITY;
            } else {
                const float eps_a = eps + 2e-5f * maxn * maxn;
                const float slack = 4.0f * (float)p.d * 6e-8f * (qn + maxn) * (qn + maxn);
                b = count <= p.cand_cap ? qn2 - 2.0f * (p.theta[qi] + eps_a) - slack : -INFINITY;
            }
            p.bound[qi] = b;
        }
    }
}

// ---- single-query pre-filter (option `prefilter`): lists of the approximate scan -> re-rank inputs -------------
// The scan kernel has just ranked the bf16 shadow for query q: Dp/Ip[q, 0..kp) = its best kp rows by approximate
// score (row positions, best-first, padded with -1).  This turns them into the candidate list, the threshold
// theta[q] = the kp-th approximate score (every row outside the list scores no better; -inf when the list holds every
// valid row) and |q|^2 for the re-rank kernel's certificate.  aug != 0: the query was extended by two -1 columns (L2
// through q.y - |y|^2/2); they do not count towards |q|^2.  One CTA per query.
__global__ void __launch_bounds__(256)
prefilter_lists_kernel(const float* __restrict__ Dp, const int64_t* __restrict__ Ip, int kp, const float* __restrict__ q, int d, int qld,
                       uint32_t* __restrict__ cand_rows, unsigned int* __restrict__ cand_count, float* __restrict__ theta,
                       float* __restrict__ qnorm2) {
    __shared__ float red[8];
    __shared__ unsigned int cnt;
    const int qi = blockIdx.x;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    unsigned int mine = 0;
    for (int j = threadIdx.x; j < kp; j += blockDim.x) {
        const int64_t r = Ip[(size_t)qi * kp + j];
        if (r >= 0) {
            cand_rows[(size_t)qi * kp + j] = (uint32_t)r;  // valid entries are a prefix of the list
            ++mine;
        }
    }
    if (mine) atomicAdd(&cnt, mine);
    float acc = 0.0f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
        const float v = q[(size_t)qi * qld + i];
        acc = fmaf(v, v, acc);
    }
    acc = warp_sum_xor(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        qnorm2[qi] = t;
        cand_count[qi] = cnt;
        theta[qi] = cnt == (unsigned)kp ? Dp[(size_t)qi * kp + kp - 1] : -INFINITY;
    }
}

// q'[nq, d + 2] = [q, -1, -1]: the extended query of the L2 pre-filter
__global__ void extend_query_kernel(const float* __restrict__ q, int64_t nq, int d, float* __restrict__ out) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * (d + 2)) return;
    const int64_t qi = t / (d + 2);
    const int c = (int)(t - qi * (d + 2));
    out[t] = c < d ? q[qi * d + c] : -1.0f;
}
