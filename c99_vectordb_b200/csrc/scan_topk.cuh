// scan_topk.cuh — K2: the single-query GEMV scan with a fused top-k (query blocks of up to 8 for
// the cases the tensor-core path does not take: masks, tiny databases, its certificate fallback).
//
// Replaces index.search(x[1,d], k) (memo_cli.py:292 -> faiss IndexFlat::search, "seq" path
// [upstream]).  HBM-bound: every database row is read exactly once per launch, nothing but the
// final [nq,k] result is written.
//
//   VARIANT_BULK : every warp owns a private ring of shared-memory stages that it fills itself
//                  with 1-D cp.async.bulk (TMA engine, SASS UBLKCP) and consumes with 128-bit
//                  LDS.  No block-level synchronisation in the main loop.
//   VARIANT_LDG  : same arithmetic, rows read with ld.global.nc.L1::no_allocate.v4 straight
//                  into registers.  Universal fallback (any row pitch) and the A/B partner for
//                  the ncu evidence.
//
// Both produce bit-identical scores: lane l of the LPR lanes sharing a row accumulates 16-byte chunks
// l, l+LPR, ... in ascending element order with fmaf, then an xor butterfly LPR/2 ... 1 (LPR = 32, 16
// for rows of <= 48 chunks, 8 for rows of 8 / 16 chunks; oracle/flat_oracle.c: score_device restates exactly this order).
//
// Top-k: per warp an unsorted k-entry list of 64-bit keys in shared memory plus the running
// threshold tau = worst key kept; a row is inserted only when its key beats tau (rare after
// warm-up: ~k*ln(rows_per_warp/k) inserts per warp per scan).  At the end of the scan the CTA
// keeps the keys that can still matter (>= the best k-th key of its warps: that warp alone proves
// k better-or-equal candidates), writes them unsorted together with their count, and raises the
// launch's global threshold (atomicMax); the LAST CTA to finish (atomic ticket) reads every CTA's
// survivors in one pass, drops what is below the global threshold, sorts the few dozen keys that
// remain, translates row -> record id (K5) and writes D/I — so a search is a single launch and its
// tail is three L2 round trips.  On a sharded index the same tail also exchanges the result with
// the peer GPUs over NVLink and merges (exchange_and_merge).  Tiles are claimed in ascending runs
// from a global counter whose run length shrinks towards the end of the database (guided
// self-scheduling) unless p.dynamic == 0.
//
// Programmatic dependent launch (p.pdl): back-to-back searches on one stream overlap — the next
// launch's CTAs take over SMs as this launch's CTAs exit, so its ramp-up (and, when the caller
// promises stable queries, its whole scan) hides this launch's merge + exchange.  Per-launch
// control words (ticket, tile counter, thresholds, survivor lists) are double-buffered by launch
// parity; every CTA executes griddepcontrol.wait before it signals launch_dependents, so launch
// i+2 cannot start before launch i has completed.
#pragma once
#include "common.cuh"

#define B200_VARIANT_BULK 1
#define B200_VARIANT_LDG 2

#define B200_SCAN_THREADS_BULK 512  // ring variant: one persistent CTA per SM, up to 16 warps
#define B200_SCAN_THREADS_LDG 256   // direct-load variant: 4 CTAs of 8 warps per SM (<= 64 registers)
#define B200_FUSED_K_MAX 256
#define B200_FINAL_BUF_KEYS 2048
#define B200_PREF_BYTES 8192u       // one 64-bit selector key per CTA in the final merge: grids of up to 1024 CTAs

struct ScanParams {
    const uint8_t* rows;      // row storage
    uint64_t pitch_bytes;     // bytes per row (multiple of 16)
    uint32_t nvec;            // 16-byte vectors per row
    uint64_t n;               // rows
    const float* q;           // queries of this launch [nqb, d] (device)
    int d;                    // logical dimension
    int qstride;              // padded floats per query in shared memory (multiple of 8)
    int nqb;                  // queries in this launch (<= QB)
    int k;                    // results per query (top-k mode)
    uint64_t* partials;       // [grid, QB, k] per-CTA best keys, sorted best-first, zero padded
    unsigned int* ticket;     // [0] CTA-done ticket, [1] tile counter; zero before launch, reset by the last CTA
    int dynamic;              // 1: tiles claimed in order from the global counter; 0: static striding
    uint32_t claim_chunk;     // dynamic: most consecutive tiles taken per atomic claim (>= 1)
    uint32_t claim_min;       // dynamic: fewest (the run shrinks with the tiles that remain)
    uint32_t claim_first;     // dynamic: every warp's first run is static (warp w: tiles [w, w+1) * claim_first), so a
                              // launch does not start with one atomic per warp on a single address (~2.5 ns each)
    int pdl;                  // 0: plain launch; 1: launched with programmatic stream serialisation, queries may come
                              // from the preceding kernel (wait before reading them); 2: queries are stable (wait after the scan)
    int normalize_q;          // 1: L2-normalise the queries while staging them (K1 arithmetic, memo_cli.py:131-135)
    unsigned long long* stamps;  // optional [grid, 8] globaltimer stamps of the phases (null = off)
    int fused_tail;           // 1: the last CTA does the final merge; 0: final_merge_kernel follows
    float* D;                 // [nqb, k] out
    int64_t* I;               // [nqb, k] out
    const int64_t* id_map;    // row -> record id, or null
    int64_t id_base;          // added to the row when id_map is null
    uint32_t tile_rows;       // rows per tile (multiple of RB)
    uint32_t stages;          // BULK: stages per warp
    uint32_t tile_bytes;      // BULK: tile_rows * pitch_bytes
    int evict_first;          // BULK: L2 evict-first hint on the stream
    uint32_t* score_keys;     // full-rank mode: [nqb, n] hi keys, else null
    const uint32_t* row_mask; // optional bitmap over rows (bit r of word r>>5): 0 = row excluded
    // fused multi-GPU exchange (null = off): xchg_peers[g] is rank g's exchange buffer as mapped in
    // this process (CUDA IPC peer mapping over NVLink); see exchange_and_merge()
    uint8_t* const* xchg_peers;
    int xchg_world, xchg_rank;
    uint32_t xchg_epoch;      // strictly increasing per launch
    uint32_t xchg_slot_bytes; // bytes of one (parity, sender) slot
    int* xchg_status;         // set to 1 if a peer never showed up
    uint32_t scratch_keys;    // power of two >= max(warps*k, B200_FINAL_BUF_KEYS)
};

// ---- small device pieces ---------------------------------------------------------------------
template <int METRIC>
__device__ __forceinline__ void acc1(float& a, float v, float q) {
    if (METRIC == 0) {
        a = fmaf(v, q, a);
    } else {
        float t = v - q;
        a = fmaf(t, t, a);
    }
}
template <int METRIC>
__device__ __forceinline__ void acc_f32x4(float& a, const uint4& raw, const float4& q) {
    acc1<METRIC>(a, __uint_as_float(raw.x), q.x);
    acc1<METRIC>(a, __uint_as_float(raw.y), q.y);
    acc1<METRIC>(a, __uint_as_float(raw.z), q.z);
    acc1<METRIC>(a, __uint_as_float(raw.w), q.w);
}
template <int METRIC>
__device__ __forceinline__ void acc_bf16x8(float& a, const uint4& raw, const float4& qa,
                                           const float4& qb) {
    acc1<METRIC>(a, __uint_as_float(raw.x << 16), qa.x);
    acc1<METRIC>(a, __uint_as_float(raw.x & 0xffff0000u), qa.y);
    acc1<METRIC>(a, __uint_as_float(raw.y << 16), qa.z);
    acc1<METRIC>(a, __uint_as_float(raw.y & 0xffff0000u), qa.w);
    acc1<METRIC>(a, __uint_as_float(raw.z << 16), qb.x);
    acc1<METRIC>(a, __uint_as_float(raw.z & 0xffff0000u), qb.y);
    acc1<METRIC>(a, __uint_as_float(raw.w << 16), qb.z);
    acc1<METRIC>(a, __uint_as_float(raw.w & 0xffff0000u), qb.w);
}

// Replace the current worst entry of a warp's list with `key`, then recompute the worst.
__device__ __forceinline__ void warp_list_insert(uint64_t* list, int k, int lane, uint64_t key,
                                                 uint64_t& tau, int& tau_pos) {
    if (lane == 0) list[tau_pos] = key;
    __syncwarp();
    uint64_t m = ~0ull;
    int mp = 0;
    for (int i = lane; i < k; i += 32) {
        uint64_t v = list[i];
        if (v < m) {
            m = v;
            mp = i;
        }
    }
    uint64_t wm = warp_min_u64(m);
    unsigned b = __ballot_sync(B200_FULL_MASK, m == wm);
    int src = __ffs(b) - 1;
    tau = wm;
    tau_pos = __shfl_sync(B200_FULL_MASK, mp, src);
}

// In-place descending bitonic sort of m (power of two) keys in shared memory by the whole CTA.
// Pair t of a compare-exchange stage is handled by thread t % blockDim, so for strides <= 32 every
// aligned 64-key block is touched by exactly one warp: those stages only need __syncwarp(); a block
// barrier is needed only around stages with stride >= 64 (15 instead of 66 barriers for m = 2048).
__device__ __forceinline__ void cta_bitonic_sort_desc(uint64_t* a, uint32_t m) {
    __syncthreads();
    for (uint32_t size = 2; size <= m; size <<= 1) {
        for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 64) __syncthreads(); else __syncwarp();
            for (uint32_t t = threadIdx.x; t < (m >> 1); t += blockDim.x) {
                uint32_t lo = 2 * t - (t & (stride - 1));  // index with bit `stride` clear
                uint32_t hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t x = a[lo], y = a[hi];
                bool swap = desc ? (x < y) : (x > y);
                if (swap) {
                    a[lo] = y;
                    a[hi] = x;
                }
            }
            if (stride == 64) __syncthreads();  // the warp-local stages that follow read other warps' writes
        }
    }
    __syncthreads();
}

// Rank-by-counting selection in shared memory: keys[0..n) are distinct 64-bit keys (0 = empty slot).  Every non-empty
// key counts the keys above it and, if fewer than k are, writes itself to out[rank]: out[0..k) ends up sorted
// best-first without a sort network or atomics; slots past the number of non-empty keys are zeroed.  n^2 / threads
// comparisons of broadcast shared-memory reads — used for n <= B200_COUNT_SELECT_MAX.  Whole CTA; `counter`: one
// shared word.  out may be global or shared memory (not aliasing keys).
#define B200_COUNT_SELECT_MAX 1024
__device__ __forceinline__ void cta_count_select(const uint64_t* keys, uint32_t n, int k, uint64_t* out, unsigned int* counter) {
    if (threadIdx.x == 0) *counter = 0u;
    __syncthreads();
    unsigned int mine = 0;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key == 0ull) continue;
        ++mine;
        uint32_t rank = 0;
        for (uint32_t j = 0; j < n; ++j) rank += keys[j] > key ? 1u : 0u;
        if (rank < (uint32_t)k) out[rank] = key;
    }
    if (mine) atomicAdd(counter, mine);
    __syncthreads();
    for (uint32_t r = *counter + threadIdx.x; r < (uint32_t)k; r += blockDim.x) out[r] = 0ull;
    __syncthreads();  // the counter may be reused at once
}

// Final merge of query qi by the whole CTA.  Every CTA left its best k keys sorted best-first (zero padded).
//  1. Threshold: with j = ceil(k / min(k, nctas)) and m = ceil(k / j), T = the m-th largest of the CTAs' j-th keys.
//     m lists hold >= j keys >= T each, i.e. >= k keys in total, so the global k-th best cannot be below T.  For
//     k <= nctas this is the k-th largest CTA maximum — tight: typically only k..2k keys are >= T.
//  2. One pass over all nctas*k keys (the first batch is already in flight while T is computed) keeps the keys
//     >= T in `scratch` (B keys); a window of the key sequence never holds more keys than the buffer has room for, and
//     adversarial inputs (thousands of keys >= T) take further windows, re-sorting and tightening T between them.
//  3. The survivors are sorted, the first k translated row -> record id (K5) and written.
// s_sel: nctas 64-bit words of shared memory.
template <int METRIC, int QB>
__device__ __forceinline__ void final_merge_one(const ScanParams& p, int qi, uint32_t nctas, uint64_t* scratch,
                                                unsigned int* sctr, uint64_t* s_sel) {
    const int k = p.k;
    const uint32_t B = p.scratch_keys;
    const uint32_t total = nctas * (uint32_t)k;
    const uint32_t kk = (uint32_t)k < nctas ? (uint32_t)k : nctas;
    const uint32_t j = ((uint32_t)k + kk - 1) / kk, m = ((uint32_t)k + j - 1) / j;
    auto key_at = [&](uint32_t i) {  // flat index over (cta, position)
        const uint32_t c = i / (uint32_t)k, r = i - c * (uint32_t)k;
        return __ldcg(p.partials + ((size_t)c * QB + qi) * k + r);
    };
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < nctas; c += blockDim.x) s_sel[c] = __ldcg(p.partials + ((size_t)c * QB + qi) * k + (j - 1));
    constexpr int U = 8;  // keys in flight per thread
    uint64_t key[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const uint32_t i = threadIdx.x + (uint32_t)u * blockDim.x;
        key[u] = i < total ? key_at(i) : 0ull;
    }
    if (threadIdx.x == 0) sctr[0] = 0u;  // sctr[1], sctr[2]: high / low word of T
    __syncthreads();
    for (uint32_t c = threadIdx.x; c < nctas; c += blockDim.x) {
        const uint64_t v = s_sel[c];
        uint32_t rank = 0;  // entries ahead of this one: larger value, or equal value (zeros) at a smaller index
        for (uint32_t c2 = 0; c2 < nctas; ++c2) {
            const uint64_t w = s_sel[c2];
            rank += (w > v || (w == v && c2 < c)) ? 1u : 0u;
        }
        if (rank == m - 1) {
            sctr[1] = (uint32_t)(v >> 32);
            sctr[2] = (uint32_t)v;
        }
    }
    __syncthreads();
    const uint64_t T = ((uint64_t)sctr[1] << 32) | sctr[2];
    uint64_t thr = T ? T - 1ull : 0ull;  // keep key > thr  <=>  key >= T (any non-empty key when fewer than m lists reach j keys)
    uint32_t pos = 0, filled = 0;
    bool preloaded = true;  // key[] holds items threadIdx.x + u * blockDim.x
    do {
        const uint32_t room = B - filled;
        const uint32_t end = total - pos < room ? total : pos + room;
        for (uint32_t i0 = pos + threadIdx.x; i0 < end; i0 += blockDim.x * U) {
            const bool have = preloaded && i0 == threadIdx.x;
            preloaded = false;
            if (!have) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                    key[u] = i < total ? key_at(i) : 0ull;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t i = i0 + (uint32_t)u * blockDim.x;
                if (i < end && key[u] > thr) scratch[atomicAdd(sctr, 1u)] = key[u];
            }
        }
        pos = end;
        __syncthreads();
        filled = *sctr;
        uint32_t mm = 2;
        while (mm < filled) mm <<= 1;
        for (uint32_t i = filled + threadIdx.x; i < mm; i += blockDim.x) scratch[i] = 0ull;
        cta_bitonic_sort_desc(scratch, mm);
        if (filled > (uint32_t)k) filled = (uint32_t)k;
        if (pos < total) {  // more windows: the k-th best so far is the new threshold (keys are unique)
            if (filled == (uint32_t)k && scratch[k - 1] > thr) thr = scratch[k - 1];
            __syncthreads();
            if (threadIdx.x == 0) *sctr = filled;
            __syncthreads();
        }
    } while (pos < total);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = (uint32_t)i < filled ? scratch[i] : 0ull;
        float dist;
        int64_t id;
        if (key == 0ull) {
            dist = (METRIC == 0) ? -FLT_MAX : FLT_MAX;
            id = -1;
        } else {
            dist = b200_key_score(key, METRIC);
            uint32_t row = b200_key_row(key);
            id = p.id_map ? p.id_map[row] : (int64_t)row + p.id_base;
        }
        if (p.xchg_peers) {
            // The local result goes straight into slot [parity][my rank] of EVERY rank's exchange buffer as three
            // self-validating 8-byte words per entry (value, epoch) — 8-byte stores are single transactions, so the
            // receiver needs no fence and no separate flag: a word whose upper half carries this launch's epoch is
            // complete.  Posted stores over NVLink (a plain store for the own buffer).
            const size_t slot = ((size_t)(p.xchg_epoch & 1u) * p.xchg_world + p.xchg_rank) * p.xchg_slot_bytes;
            const size_t off = 16 + ((size_t)qi * k + i) * 24;
            const uint64_t ep = (uint64_t)p.xchg_epoch << 32;
            const uint64_t w0 = ep | (uint32_t)((uint64_t)id & 0xffffffffull), w1 = ep | (uint32_t)((uint64_t)id >> 32),
                           w2 = ep | __float_as_uint(dist);
            for (int gg = 0; gg < p.xchg_world; ++gg) {
                uint64_t* dst = reinterpret_cast<uint64_t*>(p.xchg_peers[(gg + p.xchg_rank) % p.xchg_world] + slot + off);
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"(w0) : "memory");
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"(w1) : "memory");
                asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 2), "l"(w2) : "memory");
            }
        } else {
            p.D[(size_t)qi * k + i] = dist;
            p.I[(size_t)qi * k + i] = id;
        }
    }
    __syncthreads();
}

// Fused multi-GPU top-k exchange, run by the last CTA of every rank's scan kernel after its local results have been
// posted into all peers' buffers: wait until every word of the world lists in the OWN buffer carries this launch's
// epoch, then merge the world best-first lists per query with the K4 rank-by-counting rule (score best-first, lower
// rank first, earlier position first) out of shared memory and write the final D/I.  Slots are double-buffered by
// epoch parity: a rank cannot be two searches ahead of a peer because each search needs every peer's entries.
// Kernels on DIFFERENT GPUs wait on one another here — never two kernels of one GPU.  A peer that does not deliver
// within ~2 s marks the launch failed: every result is padding (-1) and *xchg_status is set.
// Padding entries are recognised by their sentinel score, not by a negative id (negative record ids are legal).
// s_h: s_h_cap words of shared memory for the entries' score keys; exchanges with more entries than that (large
// k x queries x world) look the keys up in the exchange buffer itself.
template <int METRIC>
__device__ __forceinline__ void exchange_and_merge(const ScanParams& p, uint32_t* s_h, uint32_t s_h_cap) {
    const int k = p.k, G = p.xchg_world;
    const uint8_t* own = p.xchg_peers[p.xchg_rank] + (size_t)(p.xchg_epoch & 1u) * G * p.xchg_slot_bytes;
    const int total = G * p.nqb * k;
    auto words_of = [&](int t) {  // entry t enumerates (q, g, j)
        const int q = t / (G * k), r = t - q * (G * k), g = r / k, j = r - g * k;
        return reinterpret_cast<const uint64_t*>(own + (size_t)g * p.xchg_slot_bytes + 16 + ((size_t)q * k + j) * 24);
    };
    const bool in_smem = (uint32_t)total <= s_h_cap;
    auto key_of = [&](int t) -> uint32_t {  // score key of entry t (0 = padding), once every entry has arrived
        if (in_smem) return s_h[t];
        uint64_t w2;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w2) : "l"(words_of(t) + 2) : "memory");
        const float sc = __uint_as_float((uint32_t)w2);
        return b200_score_valid<METRIC>(sc) ? b200_key_hi<METRIC>(sc) : 0u;
    };
    int64_t id0 = -1;
    float sc0 = 0.0f;
    int timed_out = 0;
    const long long t0 = clock64();
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const uint64_t* w = words_of(t);
        uint64_t w0, w1, w2;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(w) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(w + 1) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w2) : "l"(w + 2) : "memory");
            if ((uint32_t)(w0 >> 32) == p.xchg_epoch && (uint32_t)(w1 >> 32) == p.xchg_epoch && (uint32_t)(w2 >> 32) == p.xchg_epoch) break;
            if (clock64() - t0 > 4000000000ll) {  // ~2 s: a peer is gone
                timed_out = 1;
                break;
            }
        }
        const float sc = __uint_as_float((uint32_t)w2);
        if (in_smem) s_h[t] = (timed_out || !b200_score_valid<METRIC>(sc)) ? 0u : b200_key_hi<METRIC>(sc);
        if (t == (int)threadIdx.x) {
            id0 = (int64_t)((w0 & 0xffffffffull) | (w1 << 32));
            sc0 = sc;
        }
    }
    const int failed = __syncthreads_or(timed_out);
    if (failed) {
        if (threadIdx.x == 0) *p.xchg_status = 1;
        for (int t = threadIdx.x; t < p.nqb * k; t += blockDim.x) {
            p.D[t] = (METRIC == 0) ? -FLT_MAX : FLT_MAX;
            p.I[t] = -1;
        }
        return;
    }
    for (int t = threadIdx.x; t < total; t += blockDim.x) {
        const int q = t / (G * k), r = t - q * (G * k), g = r / k, j = r - g * k;
        int64_t id = id0;
        float sc = sc0;
        if (t != (int)threadIdx.x) {  // entries beyond the first sweep: read them again (they have arrived)
            const uint64_t* w = words_of(t);
            uint64_t w0, w1, w2;
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(w) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(w + 1) : "memory");
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w2) : "l"(w + 2) : "memory");
            id = (int64_t)((w0 & 0xffffffffull) | (w1 << 32));
            sc = __uint_as_float((uint32_t)w2);
        }
        const uint32_t h = key_of(t);
        int rank = j;
        for (int g2 = 0; g2 < G; ++g2) {
            if (g2 == g) continue;
            const int base2 = (q * G + g2) * k;
            int lo = 0, hi = k;  // entries of list g2 that precede this candidate
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const uint32_t hm = key_of(base2 + mid);
                const bool before = (g2 < g) ? (hm >= h) : (hm > h);
                if (before) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            p.D[(size_t)q * k + rank] = h == 0u ? ((METRIC == 0) ? -FLT_MAX : FLT_MAX) : sc;
            p.I[(size_t)q * k + rank] = h == 0u ? (int64_t)-1 : id;
        }
    }
}

// One CTA per query: the final merge as its own launch, used when a scan launch carries several
// queries (the fused last-CTA tail would merge them one after another).
template <int METRIC, int QB>
__global__ void __launch_bounds__(256) final_merge_kernel(const ScanParams p, uint32_t nctas) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* scratch = reinterpret_cast<uint64_t*>(smem);
    unsigned int* sctr = reinterpret_cast<unsigned int*>(smem + (size_t)p.scratch_keys * 8);
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(sctr + 4);
    final_merge_one<METRIC, QB>(p, blockIdx.x, nctas, scratch, sctr, s_sel);
}

// ---- the kernel --------------------------------------------------------------------------------
// LPR = lanes that share one row: 32 normally; 16 for rows of <= 48 sixteen-byte chunks (<= 768 B) and 8 for rows of
// 8 / 16 chunks (128 / 256 B; SETS = 4 row sets, butterfly 4,2,1), where a
// full warp per row would leave lanes idle — the warp then works on SETS = 2 row sets of RB rows at once
// (lanes 0-15 on the first, 16-31 on the second) and the butterfly has 4 rounds (8,4,2,1).
template <int METRIC, int STORE, int QB, int RB, int VARIANT, int LPR>
__global__ void __launch_bounds__(VARIANT == B200_VARIANT_BULK ? (QB >= 4 ? 256 : B200_SCAN_THREADS_BULK) : B200_SCAN_THREADS_LDG,
                                  VARIANT == B200_VARIANT_BULK ? 1 : (QB >= 4 ? 2 : 4))  // QB*RB accumulators need registers
scan_topk_kernel(const ScanParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int SETS = 32 / LPR;      // row sets a warp processes concurrently
    constexpr int RSTEP = RB * SETS;    // rows per warp step
    const int lane = threadIdx.x & 31;
    const int set = lane / LPR, sl = lane % LPR;
    const int warp = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const bool fullrank = (p.score_keys != nullptr);
    const int k = p.k;

    // ---- shared memory carve-up (host computes the same sizes: scan_smem_bytes) ----
    // [ring: nw*stages*tile_bytes | scratch (aliases ring start)] [queries] [lists] [mbarriers] [ctr] [stage tiles]
    uint32_t ring_bytes = (VARIANT == B200_VARIANT_BULK) ? nw * p.stages * p.tile_bytes : 0u;
    uint32_t scratch_bytes = p.scratch_keys * 8u + B200_PREF_BYTES;  // final-merge buffer + one selector key per CTA
    uint32_t region0 = ring_bytes > scratch_bytes ? ring_bytes : scratch_bytes;
    region0 = (region0 + 127u) & ~127u;
    uint8_t* ring = smem;
    uint64_t* scratch = reinterpret_cast<uint64_t*>(smem);
    uint64_t* s_sel = reinterpret_cast<uint64_t*>(smem + p.scratch_keys * 8u);
    float* qs = reinterpret_cast<float*>(smem + region0);
    uint32_t q_bytes = (uint32_t)QB * p.qstride * 4u;
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem + region0 + q_bytes);
    uint32_t list_bytes = fullrank ? 0u : (uint32_t)nw * QB * k * 8u;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + region0 + q_bytes + list_bytes);
    unsigned int* sctr = reinterpret_cast<unsigned int*>(bars + nw * (VARIANT == B200_VARIANT_BULK ? p.stages : 0u));
    uint32_t* stage_tile = reinterpret_cast<uint32_t*>(sctr + 16) + warp * (VARIANT == B200_VARIANT_BULK ? p.stages : 0u);
    __shared__ unsigned int s_is_last;

    if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 0] = globaltimer_ns();
    // ---- clear lists (the queries are staged below, after the first tiles have been requested) ----
    if (!fullrank)
        for (int i = threadIdx.x; i < nw * QB * k; i += blockDim.x) lists[i] = 0ull;

    const uint32_t TR = p.tile_rows;
    const uint32_t tiles_total = (uint32_t)((p.n + TR - 1) / TR);
    const uint32_t NOTILE = 0xFFFFFFFFu;

    // Tile scheduler.  dynamic: tiles are claimed in ascending order from one global counter, so
    // every SM streams until the database is exhausted (no tail where the slow SMs finish alone)
    // and DRAM sees one advancing front.  static: warp gw takes tiles gw, gw+GW, ...
    uint32_t static_next = blockIdx.x * nw + warp;
    const uint32_t static_step = gridDim.x * nw;
    uint32_t chunk_next = (blockIdx.x * nw + warp) * p.claim_first;  // lane 0: the claimed run of consecutive tiles
    uint32_t chunk_end = chunk_next + p.claim_first;
    const uint32_t claim_base = gridDim.x * nw * p.claim_first;       // the counter hands out tiles from here on
    const uint32_t guide_div = 2u * gridDim.x * (uint32_t)nw;
    auto claim = [&]() -> uint32_t {  // called by lane 0 only
        if (p.dynamic) {
            if (chunk_next == chunk_end) {
                // One atomic per run of tiles (a single hot address serialises at ~2.5 ns/op).  Guided
                // self-scheduling: the run is 1/(2 x warps) of what was left at this warp's previous claim,
                // between claim_min and claim_chunk tiles, so the warps run out of work within claim_min tiles
                // of one another instead of claim_chunk.
                const uint32_t left = chunk_end < tiles_total ? tiles_total - chunk_end : 0u;
                uint32_t run = left / guide_div;
                run = run < p.claim_min ? p.claim_min : (run > p.claim_chunk ? p.claim_chunk : run);
                chunk_next = claim_base + atomicAdd(p.ticket + 1, run);
                chunk_end = chunk_next + run;
            }
            return chunk_next++;
        }
        uint32_t t = static_next;
        static_next = (t > NOTILE - static_step) ? NOTILE : t + static_step;
        return t;
    };

    uint32_t bar0 = 0, ring0 = 0;
    uint64_t policy = 0;
    if (VARIANT == B200_VARIANT_BULK) {
        bar0 = smem_u32(bars + warp * p.stages);
        ring0 = smem_u32(ring + (size_t)warp * p.stages * p.tile_bytes);
        if (lane == 0) {
            for (uint32_t s = 0; s < p.stages; ++s) mbar_init(bar0 + 8u * s, 1u);
            mbar_fence_init();
        }
        if (p.evict_first) policy = l2_policy_evict_first();
    }
    __syncthreads();

    auto issue = [&](uint32_t s, uint32_t t) {
        uint64_t row0 = (uint64_t)t * TR;
        uint64_t nr = p.n - row0 < TR ? p.n - row0 : TR;
        uint32_t bytes = (uint32_t)(nr * p.pitch_bytes);
        uint32_t bar = bar0 + 8u * s;
        mbar_arrive_expect_tx(bar, bytes);
        if (p.evict_first)
            bulk_g2s_hint(ring0 + s * p.tile_bytes, p.rows + row0 * p.pitch_bytes, bytes, bar, policy);
        else
            bulk_g2s(ring0 + s * p.tile_bytes, p.rows + row0 * p.pitch_bytes, bytes, bar);
    };

    uint32_t pending = NOTILE;  // lane 0: the tile claimed one step ahead (hides the atomic's latency)
    uint32_t t_cur = NOTILE;    // LDG: the tile this warp works on
    if (VARIANT == B200_VARIANT_BULK) {
        if (lane == 0) {
            for (uint32_t s = 0; s < p.stages; ++s) {
                uint32_t t = claim();
                if (t < tiles_total) {
                    stage_tile[s] = t;
                    issue(s, t);
                } else {
                    stage_tile[s] = NOTILE;
                }
            }
            pending = claim();
        }
        __syncwarp();
    } else {
        if (lane == 0) {
            t_cur = claim();
            pending = claim();
        }
        t_cur = __shfl_sync(B200_FULL_MASK, t_cur, 0);
    }

    // ---- stage the queries (zero padded; optionally L2-normalised with K1's arithmetic) ----
    // With programmatic dependent launch everything above (and the first tiles in flight) overlapped the
    // previous kernel's tail; the queries may be its output unless the caller promised otherwise.
    if (p.pdl == 1) {
        griddep_wait();
        griddep_launch_dependents();
    }
    if (p.normalize_q) {
        const int nchunk = (p.d + 3) >> 2;
        for (int qi = warp; qi < QB; qi += nw) {
            float* dst = qs + (size_t)qi * p.qstride;
            if (qi >= p.nqb) {
                for (int c = lane; c < p.qstride; c += 32) dst[c] = 0.0f;
                continue;
            }
            const float* src = p.q + (size_t)qi * p.d;
            float acc = 0.0f;  // lane l owns the 4-element chunks l, l+32, ...: ingest_rows_kernel's order
            for (int c = lane; c < nchunk; c += 32)
                for (int e = 4 * c; e < 4 * c + 4 && e < p.d; ++e) {
                    const float v = src[e];
                    acc = fmaf(v, v, acc);
                }
            acc = warp_sum_xor(acc);
            const float nrm = __fsqrt_rn(acc);
            const bool zero = ((double)nrm <= 1e-8);  // memo_cli.py:133 compares against the double 1e-8
            for (int c = lane; c < p.qstride; c += 32) dst[c] = (c < p.d && !zero) ? __fdiv_rn(src[c], nrm) : 0.0f;
        }
    } else {
        for (int i = threadIdx.x; i < QB * p.qstride; i += blockDim.x) {
            int qi = i / p.qstride, c = i - qi * p.qstride;
            qs[i] = (qi < p.nqb && c < p.d) ? p.q[(size_t)qi * p.d + c] : 0.0f;
        }
    }
    __syncthreads();
    if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 1] = globaltimer_ns();

    uint64_t tau[QB];
    int tau_pos[QB];
#pragma unroll
    for (int qi = 0; qi < QB; ++qi) {
        tau[qi] = 0ull;
        tau_pos[qi] = 0;
    }
    uint64_t* my_lists = lists + (size_t)warp * QB * k;
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    const int qstride4 = p.qstride >> 2;
    const uint32_t nvec = p.nvec;

    for (uint32_t it = 0;; ++it) {
        uint32_t s = 0;
        const uint8_t* tile_smem = nullptr;
        uint32_t t;
        if (VARIANT == B200_VARIANT_BULK) {
            s = it % p.stages;
            t = stage_tile[s];
            if (t == NOTILE) break;
            uint32_t parity = (it / p.stages) & 1u;
            mbar_wait(bar0 + 8u * s, parity);
            if (p.stamps && it == 0 && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 7] = globaltimer_ns();
            tile_smem = ring + ((size_t)warp * p.stages + s) * p.tile_bytes;
        } else {
            t = t_cur;
            if (t >= tiles_total) break;
        }
        const uint64_t tile_row0 = (uint64_t)t * TR;
        const uint32_t rows_in_tile = (uint32_t)(p.n - tile_row0 < TR ? p.n - tile_row0 : TR);
        for (uint32_t g = 0; g < rows_in_tile; g += RSTEP) {
            float acc[QB][RB];
#pragma unroll
            for (int qi = 0; qi < QB; ++qi)
#pragma unroll
                for (int r = 0; r < RB; ++r) acc[qi][r] = 0.0f;

            const uint32_t g_set = g + (uint32_t)set * RB;  // first row (within the tile) of this lane's set
            const uint8_t* rp[RB];
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                if (VARIANT == B200_VARIANT_BULK) {
                    rp[r] = tile_smem + (size_t)(g_set + r) * p.pitch_bytes;  // stale rows are discarded below
                } else {
                    uint64_t row = tile_row0 + g_set + r;
                    if (row >= p.n) row = p.n - 1;
                    rp[r] = p.rows + row * p.pitch_bytes;
                }
            }
#pragma unroll 2
            for (uint32_t c = sl; c < nvec; c += LPR) {
                uint4 raw[RB];
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    if (VARIANT == B200_VARIANT_BULK)
                        raw[r] = *reinterpret_cast<const uint4*>(rp[r] + (size_t)c * 16);
                    else
                        raw[r] = ldg_nc_v4(rp[r] + (size_t)c * 16);
                }
                if (STORE == 0) {
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) {
                        float4 qv = q4[qi * qstride4 + c];
#pragma unroll
                        for (int r = 0; r < RB; ++r) acc_f32x4<METRIC>(acc[qi][r], raw[r], qv);
                    }
                } else {
#pragma unroll
                    for (int qi = 0; qi < QB; ++qi) {
                        float4 qa = q4[qi * qstride4 + 2 * c];
                        float4 qb = q4[qi * qstride4 + 2 * c + 1];
#pragma unroll
                        for (int r = 0; r < RB; ++r) acc_bf16x8<METRIC>(acc[qi][r], raw[r], qa, qb);
                    }
                }
            }
            if (QB == 1 && RB == 4 && !fullrank) {
                // Transposing butterfly over the LPR lanes of a row set: after the first two rounds
                // (xor LPR/2, xor LPR/4) each lane carries ONE row's partial sum, the remaining rounds
                // finish it.  The additions are the same pairs as four separate butterflies, so scores
                // stay bit-identical; 6 (5) shuffles instead of 20 (16), and one ballot decides whether
                // any of the warp's RSTEP rows beats the threshold.
                constexpr int H = LPR / 2, Q4 = LPR / 4;
                const bool upH = (sl & H) != 0, upQ = (sl & Q4) != 0;
                float k0 = upH ? acc[0][2] : acc[0][0], k1 = upH ? acc[0][3] : acc[0][1];
                float s0 = upH ? acc[0][0] : acc[0][2], s1 = upH ? acc[0][1] : acc[0][3];
                k0 += __shfl_xor_sync(B200_FULL_MASK, s0, H);
                k1 += __shfl_xor_sync(B200_FULL_MASK, s1, H);
                float kk = upQ ? k1 : k0, ss = upQ ? k0 : k1;
                kk += __shfl_xor_sync(B200_FULL_MASK, ss, Q4);
#pragma unroll
                for (int m = Q4 / 2; m >= 1; m >>= 1) kk += __shfl_xor_sync(B200_FULL_MASK, kk, m);
                const uint32_t myr = (upH ? 2u : 0u) + (upQ ? 1u : 0u);
                const uint64_t myrow = tile_row0 + g_set + myr;
                bool live = (g_set + myr < rows_in_tile) && b200_score_valid<METRIC>(kk);
                if (p.row_mask && live) live = (__ldg(p.row_mask + (myrow >> 5)) >> (myrow & 31)) & 1u;
                const uint64_t mykey = b200_make_key<METRIC>(kk, (uint32_t)myrow);
                unsigned hits = __ballot_sync(B200_FULL_MASK, live && mykey > tau[0]);
                while (hits) {  // rare: a row beats the warp's current k-th best
                    const int src = __ffs(hits) - 1;
                    const uint64_t key = __shfl_sync(B200_FULL_MASK, mykey, src);
                    hits &= ~(((1u << Q4) - 1u) << (src & ~(Q4 - 1)));  // the Q4 lanes of that row carry the same key
                    if (key > tau[0]) warp_list_insert(my_lists, k, lane, key, tau[0], tau_pos[0]);
                }
                continue;
            }
#pragma unroll
            for (int qi = 0; qi < QB; ++qi)
#pragma unroll
                for (int r = 0; r < RB; ++r) {
                    float v = acc[qi][r];
#pragma unroll
                    for (int m = LPR / 2; m >= 1; m >>= 1) v += __shfl_xor_sync(B200_FULL_MASK, v, m);
                    acc[qi][r] = v;  // every lane of a row set now holds that set's score
                }

#pragma unroll
            for (int st = 0; st < SETS; ++st) {  // warp-uniform walk over the row sets
#pragma unroll
                for (int qi = 0; qi < QB; ++qi) {
                    if (qi >= p.nqb) break;
#pragma unroll
                    for (int r = 0; r < RB; ++r) {
                        if (g + st * RB + r >= rows_in_tile) break;
                        const uint64_t row = tile_row0 + g + st * RB + r;
                        const float sc = SETS == 1 ? acc[qi][r] : __shfl_sync(B200_FULL_MASK, acc[qi][r], st * LPR);
                        bool valid = b200_score_valid<METRIC>(sc);
                        if (p.row_mask && valid) valid = (__ldg(p.row_mask + (row >> 5)) >> (row & 31)) & 1u;
                        if (fullrank) {
                            if (lane == 0)
                                p.score_keys[(size_t)qi * p.n + row] = valid ? b200_key_hi<METRIC>(sc) : 0u;
                        } else if (valid) {
                            uint64_t key = b200_make_key<METRIC>(sc, (uint32_t)row);
                            if (key > tau[qi])
                                warp_list_insert(my_lists + (size_t)qi * k, k, lane, key, tau[qi], tau_pos[qi]);
                        }
                    }
                }
            }
        }
        if (VARIANT == B200_VARIANT_BULK) {
            __syncwarp();  // every lane has finished reading this stage
            if (lane == 0) {
                uint32_t tn = pending;
                if (tn < tiles_total) {
                    stage_tile[s] = tn;
                    issue(s, tn);
                    pending = claim();
                } else {
                    stage_tile[s] = NOTILE;
                }
            }
            __syncwarp();
        } else {
            uint32_t tn = pending;
            if (lane == 0 && tn < tiles_total) pending = claim();
            t_cur = __shfl_sync(B200_FULL_MASK, tn, 0);
        }
    }

    if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 2] = globaltimer_ns();
    // ---- CTA reduction: the warps' lists -> this CTA's best k per query, sorted best-first ----
    __syncthreads();  // all warps done; every issued bulk copy has been consumed
    if (p.pdl == 2) {  // stable queries: this launch only has to be ordered behind the previous one from here on
        griddep_wait();
        griddep_launch_dependents();
    }
    if (!fullrank) {
        const uint32_t nkeys = (uint32_t)(nw * k);
        for (int qi = 0; qi < p.nqb; ++qi) {
            uint64_t* out = p.partials + ((size_t)blockIdx.x * QB + qi) * k;
            if (nkeys <= B200_COUNT_SELECT_MAX && (QB == 1 || nw == 1)) {
                // the warps' lists of one query are contiguous: rank every key by counting (no sort, no atomics)
                cta_count_select(lists + (size_t)qi * k, nkeys, k, out, sctr);
            } else {
                uint32_t mm = 2;
                while (mm < nkeys) mm <<= 1;
                for (uint32_t i = threadIdx.x; i < mm; i += blockDim.x) {
                    uint64_t v = 0ull;
                    if (i < nkeys) {
                        uint32_t w = i / k, jj = i - w * k;
                        v = lists[((size_t)w * QB + qi) * k + jj];
                    }
                    scratch[i] = v;
                }
                if (mm <= B200_COUNT_SELECT_MAX) {
                    __syncthreads();
                    cta_count_select(scratch, nkeys, k, out, sctr);
                } else {
                    cta_bitonic_sort_desc(scratch, mm);
                    for (int i = threadIdx.x; i < k; i += blockDim.x) out[i] = scratch[i];
                }
                __syncthreads();
            }
        }
    }
    if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 3] = globaltimer_ns();

    // ---- last CTA: merge all survivors, translate ids, write D/I ----
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(p.ticket, 1u);
        s_is_last = (prev == gridDim.x - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 4] = globaltimer_ns();
    if (!fullrank && p.fused_tail) {
        for (int qi = 0; qi < p.nqb; ++qi) final_merge_one<METRIC, QB>(p, qi, gridDim.x, scratch, sctr, s_sel);
        if (p.stamps && threadIdx.x == 0) p.stamps[blockIdx.x * 8 + 5] = globaltimer_ns();
        if (p.xchg_peers) exchange_and_merge<METRIC>(p, reinterpret_cast<uint32_t*>(scratch), p.scratch_keys * 2u);
    }
    if (threadIdx.x == 0) {  // ready for the launch after next (this parity set)
        p.ticket[0] = 0u;
        p.ticket[1] = 0u;
        if (p.stamps) p.stamps[blockIdx.x * 8 + 6] = globaltimer_ns();
    }
}

// shared memory the kernel needs for a configuration (host side twin of the carve-up above)
static inline size_t scan_smem_bytes(int variant, int nw, int QB, int qstride, int k, bool fullrank,
                                     uint32_t stages, uint32_t tile_bytes, uint32_t scratch_keys) {
    size_t ring = variant == B200_VARIANT_BULK ? (size_t)nw * stages * tile_bytes : 0;
    size_t scratch = (size_t)scratch_keys * 8 + B200_PREF_BYTES;
    size_t region0 = ring > scratch ? ring : scratch;
    region0 = (region0 + 127) & ~(size_t)127;
    size_t q = (size_t)QB * qstride * 4;
    size_t lists = fullrank ? 0 : (size_t)nw * QB * k * 8;
    size_t bars = variant == B200_VARIANT_BULK ? (size_t)nw * stages * (8 + 4) : 0;  // mbarriers + stage tiles
    return region0 + q + lists + bars + 16 + 4 * 16;  // + the per-query survivor counters
}
