// gemm_ts.cuh — K3, queries resident in TENSOR MEMORY (the "TS" form of tcgen05.mma: A from TMEM, B from smem).
//
// The 256 x 256 kernel of gemm_topk.cuh re-fetches the 128-query operand tile of every CTA for every 256-row tile:
// per 64-column K block an SM receives 32 KB of TMA fills and its tensor core reads 32 KB of operands — the whole
// 128 B/cycle shared-memory port, and twice the L2 -> SM traffic the rows alone would need (ncu: tensor pipe 77 %
// active with the MMA warp never waiting).  A work unit, however, sweeps `chunk_tiles` row tiles with the SAME 256
// queries, so here they stay put:
//   * each CTA of the pair keeps its 128 queries as the A operand in tensor memory — lane = query, 32-bit column c =
//     K elements (2c, 2c+1), i.e. the bf16 row as it lies in memory (kpad / 2 columns, written once per work unit
//     by the epilogue warps with tcgen05.st, one query per thread);
//   * only the database rows stream: NT-row tiles, each CTA stages NT/2 rows per K block through a TMA ring;
//   * the rest of tensor memory holds two NT-column accumulators (double buffered): kpad <= 512 -> NT = 128,
//     kpad <= 768 -> NT = 64 (384 + 2 x 64 = 512 columns at d = 768); longer rows keep the 256 x 256 kernel.
// Shared-memory traffic per SM falls from 64 KB to 16 KB (NT = 64: 8 KB) per 64 K columns of a 256-row sweep and the
// L2 -> SM traffic halves.  Epilogue, candidate lists, thresholds, re-rank and certificate are those of gemm_topk.cuh;
// a "tile" on the host side stays 256 rows (sampling, tilemax layout), the kernel walks its 256 / NT sub-tiles.
#pragma once
#include "gemm_topk.cuh"

#define G3S_THREADS 192
#define G3S_KB_BYTES(NT) ((NT) / 2 * G3_BLOCK_K * 2)  // one CTA's half of an NT-row tile, one K block

__device__ __forceinline__ void tc_st_32x32b_x32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
          "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
          "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem desc]: cta_group::2, M = 256 (128 lanes per CTA), K = 16; issued by the elected lane
// (warp-uniform issue, see elect_one_pred in gemm_topk.cuh)
__device__ __forceinline__ void tc_mma_f16_ts_cg2_if(uint32_t elected, uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "setp.ne.b32 e, %5, 0;\n\t"
        "@e tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(elected)
        : "memory");
}
struct GemmTsParams {
    GemmParams g;           // tiles are 256 rows, as for gemm_topk_kernel
    const uint32_t* qb;     // bf16 query shadow [m_tiles rounded up to pairs * 128, kpad], zero padded, read as 32-bit words
    uint32_t kps;           // K blocks per ring stage (divides k_blocks)
    uint32_t stages;        // ring stages
};

template <int NT, bool MASKED, int MODE>
__global__ void __launch_bounds__(G3S_THREADS, 1)
gemm_topk_ts_kernel(const __grid_constant__ CUtensorMap tm_db, const GemmTsParams tp) {
    const GemmParams& p = tp.g;
    constexpr uint32_t kKbBytes = G3S_KB_BYTES(NT);
    constexpr uint32_t kSub = G3_BLOCK_N / NT;  // sub-tiles per 256-row tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t S = tp.stages, kps = tp.kps;
    const uint32_t stage_bytes = kps * kKbBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S * stage_bytes);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty, [2S+4]: queries in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 5);
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (S + s); };
    auto tfull_bar = [&](uint32_t b) { return bar_base + 8u * (2 * S + b); };
    auto tempty_bar = [&](uint32_t b) { return bar_base + 8u * (2 * S + 2 + b); };
    const uint32_t aready_bar = bar_base + 8u * (2 * S + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t group_id = blockIdx.x / 2, groups = gridDim.x / 2;
    const uint32_t a_cols = p.k_blocks * (G3_BLOCK_K / 2);  // 32-bit columns of the resident queries
    const uint32_t acc_col0 = 512u - 2u * NT;               // the two accumulators sit at the top of tensor memory

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_db);
        for (uint32_t s = 0; s < S; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (uint32_t b = 0; b < 2; ++b) {
            mbar_init(tfull_bar(b), 1);
            mbar_init(tempty_bar(b), 8);  // four epilogue warps of each CTA of the pair
        }
        mbar_init(aready_bar, 8);
        mbar_fence_init();
    }
    if (warp == 1) tc_alloc_cg<2>(smem_u32(tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t chunks = (p.tile_count + p.chunk_tiles - 1) / p.chunk_tiles;
    const uint32_t m_groups = (p.m_tiles + 1) / 2;  // 256 queries per work unit
    const uint32_t units = chunks * m_groups;       // unit u -> (chunk = u / m_groups, m group = u % m_groups)

    if (warp == 0) {
        // ===== TMA producer: the rows only =====
        {
            const uint32_t elected = elect_one_pred();
            uint32_t stage = 0, phase = 0;
            for (uint32_t u = group_id; u < units; u += groups) {
                const uint32_t chunk = u / m_groups;
                const uint32_t t0 = chunk * p.chunk_tiles;
                const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
                for (uint32_t t = t0; t < t1; ++t) {
                    const uint32_t src_row0 = (p.src_tile_first + t * p.src_tile_stride) * G3_BLOCK_N + rank * (NT / 2);
                    for (uint32_t sub = 0; sub < kSub; ++sub) {
                        for (uint32_t kb = 0; kb < p.k_blocks; kb += kps) {
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            if (leader) mbar_arrive_expect_tx_if(elected, full_bar(stage), stage_bytes * 2);
                            const uint32_t lead_bar = mapa_shared(full_bar(stage), 0);
                            const uint32_t dst = smem_base + stage * stage_bytes;
                            for (uint32_t j = 0; j < kps; ++j)
                                tma_load_2d_cg2_if(elected, dst + j * kKbBytes, &tm_db, (int)((kb + j) * G3_BLOCK_K), (int)(src_row0 + sub * NT), lead_bar);
                            if (++stage == S) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (the pair's leader) =====
        if (leader) {
            const uint32_t idesc = umma_idesc_bf16(256, NT);
            const uint32_t elected = elect_one_pred();
            uint32_t stage = 0, phase = 0, abuf = 0, aphase = 0, qphase = 0;
            for (uint32_t u = group_id; u < units; u += groups) {
                const uint32_t chunk = u / m_groups;
                const uint32_t t0 = chunk * p.chunk_tiles;
                const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
                mbar_wait(aready_bar, qphase);  // this unit's queries are in tensor memory (both CTAs)
                qphase ^= 1u;
                tc_fence_after();
                for (uint32_t ts = t0 * kSub; ts < t1 * kSub; ++ts) {
                    mbar_wait(tempty_bar(abuf), aphase ^ 1u);
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + acc_col0 + abuf * NT;
                    for (uint32_t kb = 0; kb < p.k_blocks; kb += kps) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t b_addr = smem_base + stage * stage_bytes;
                        for (uint32_t j = 0; j < kps; ++j) {
                            const uint64_t bdesc = umma_desc_k_sw128(b_addr + j * kKbBytes);
                            const uint32_t a_tmem = tmem_base + (kb + j) * (G3_BLOCK_K / 2);
#pragma unroll
                            for (uint32_t k = 0; k < G3_BLOCK_K / G3_UMMA_K; ++k)
                                tc_mma_f16_ts_cg2_if(elected, d_tmem, a_tmem + k * (G3_UMMA_K / 2), bdesc + 2 * k, idesc, (kb | j | k) != 0 ? 1u : 0u);
                        }
                        tc_commit_cg_if<2>(elected, empty_bar(stage));
                        if (kb + kps >= p.k_blocks) tc_commit_cg_if<2>(elected, tfull_bar(abuf));
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
        }
    } else {
        // ===== epilogue warps: own 32 TMEM lanes each — one query per thread =====
        const uint32_t quarter = (uint32_t)warp & 3u;
        const uint32_t lane_base = tmem_base + ((quarter * 32u) << 16);
        const uint32_t tempty_lead0 = mapa_shared(tempty_bar(0), 0);
        const uint32_t aready_lead = mapa_shared(aready_bar, 0);
        // the unit's queries -> tensor memory: thread = query, 32 columns (one 128-byte K block of the bf16 row) per store
        auto load_queries = [&](uint32_t mt) {
            const uint32_t* src = tp.qb + ((size_t)mt * G3_BLOCK_M + quarter * 32u + (uint32_t)lane) * a_cols;
            for (uint32_t kb = 0; kb < p.k_blocks; ++kb) {
                uint32_t v[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 w = __ldg(reinterpret_cast<const uint4*>(src + kb * 32u) + i);
                    v[4 * i] = w.x; v[4 * i + 1] = w.y; v[4 * i + 2] = w.z; v[4 * i + 3] = w.w;
                }
                tc_st_32x32b_x32(lane_base + kb * 32u, v);
            }
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(aready_lead);
        };
        uint32_t abuf = 0, aphase = 0;
        bool first = true;
        for (uint32_t u = group_id; u < units; u += groups) {
            const uint32_t chunk = u / m_groups, mg = u - chunk * m_groups;
            const uint32_t mt = mg * 2 + rank;
            const uint32_t t0 = chunk * p.chunk_tiles;
            const uint32_t t1 = min(t0 + p.chunk_tiles, p.tile_count);
            const uint32_t qidx = mt * G3_BLOCK_M + quarter * 32u + (uint32_t)lane;
            const bool qlive = qidx < p.nq;
            float theta = INFINITY;
            if (MODE == G3_MODE_EMIT && qlive) theta = p.theta[qidx];
            // every MMA of the previous unit has retired: its last accumulator was waited for below
            if (first) load_queries(mt);
            first = false;
            for (uint32_t t = t0; t < t1; ++t) {
                const uint32_t ntile = p.tile_first + t * p.tile_stride;
                float gmax[8];
#pragma unroll
                for (uint32_t sub = 0; sub < kSub; ++sub) {
                    const uint64_t row0 = (uint64_t)ntile * G3_BLOCK_N + sub * NT;
                    const uint32_t live_cols = row0 >= p.n ? 0u : (uint32_t)(p.n - row0 < NT ? p.n - row0 : NT);
                    mbar_wait(tfull_bar(abuf), aphase);
                    tc_fence_after();
                    const uint32_t taddr0 = lane_base + acc_col0 + abuf * NT;
                    auto group = [&](const uint32_t (&v)[32], uint32_t g) {
                        const uint32_t c0 = g * 32u;
                        uint32_t ok = c0 >= live_cols ? 0u : (live_cols - c0 >= 32u ? 0xFFFFFFFFu : (1u << (live_cols - c0)) - 1u);
                        if (MASKED && ok) ok &= __ldg(p.row_mask + ((row0 + c0) >> 5));
                        float m;
                        if (ok == 0xFFFFFFFFu) m = g3_max32(v);
                        else if (ok == 0u) m = -INFINITY;
                        else m = g3_max32_where(v, ok);
                        if (MODE == G3_MODE_TILEMAX) gmax[sub * (NT / 32) + g] = m;
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
                    uint32_t va[32], vb[32];
                    tc_ld_32x32b_x32(taddr0, va);
#pragma unroll
                    for (uint32_t g = 0; g < NT / 32; g += 2) {
                        tc_wait_ld_for(va);
                        tc_ld_32x32b_x32(taddr0 + (g + 1) * 32u, vb);
                        group(va, g);
                        tc_wait_ld_for(vb);
                        if (g + 2 < NT / 32) {
                            tc_ld_32x32b_x32(taddr0 + (g + 2) * 32u, va);
                        } else {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive_cluster(tempty_lead0 + 8u * abuf);
                        }
                        group(vb, g + 1);
                    }
                    if (++abuf == 2) {
                        abuf = 0;
                        aphase ^= 1u;
                    }
                }
                if (MODE == G3_MODE_TILEMAX && qlive) {
                    float4* dst = reinterpret_cast<float4*>(p.tilemax + ((size_t)qidx * p.tile_count + t) * 8);
                    dst[0] = make_float4(gmax[0], gmax[1], gmax[2], gmax[3]);
                    dst[1] = make_float4(gmax[4], gmax[5], gmax[6], gmax[7]);
                }
            }
            // the last accumulator of this unit has been read, so every MMA that used this unit's queries is done:
            // replace them with the next unit's
            const uint32_t un = u + groups;
            if (un < units) load_queries((un % m_groups) * 2 + rank);
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tc_dealloc_cg<2>(tmem_base, 512);
    }
}
