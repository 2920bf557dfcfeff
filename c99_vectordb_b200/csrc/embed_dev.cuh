// embed_dev.cuh — K6: the hashing-trick embedder on the device, fused with K1.
//
// Replaces the per-record Python loop of rebuild_index_from_texts (memo_cli.py:272-285): for every non-blank
// record, embed_text_hash (memo_cli.py:158-167) = tokens [a-z0-9_]+ of the lower-cased text, h = hash(token),
// vec[|h| % dim] += (h & 1) ? +1 : -1, then normalize() (memo_cli.py:131-135) — and add_with_ids of the row.
// Here the text bytes are the only thing that crosses PCIe: one warp per record tokenises its bytes with ballots,
// every token start lane hashes its token (CPython's SipHash-1-3 under the PYTHONHASHSEED=0 key, as csrc/embed.cu),
// buckets accumulate in shared memory (+-1 sums are exact in fp32, so the order of the atomics does not matter),
// and the row is normalised with K1's arithmetic and stored straight into the index (fp32 or bf16): text -> resident
// row in one kernel.  Blank records (only ASCII white space; memo_cli.py:141-142, :277-279) are skipped; kept rows
// stay in record order (row position monotone in record id, which the tie rule relies on).
#pragma once
#include "common.cuh"

#define EMB_BLOCK_RECS 256   // records per CTA
#define EMB_THREADS 256      // 8 warps
#define EMB_MAX_DIM 2048     // bucket vector per warp in shared memory: 8 warps x 8 KB

__device__ __forceinline__ uint64_t emb_rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
#define EMB_SIPROUND                                                      \
    do {                                                                  \
        v0 += v1; v1 = emb_rotl(v1, 13); v1 ^= v0; v0 = emb_rotl(v0, 32); \
        v2 += v3; v3 = emb_rotl(v3, 16); v3 ^= v2;                        \
        v0 += v3; v3 = emb_rotl(v3, 21); v3 ^= v0;                        \
        v2 += v1; v1 = emb_rotl(v1, 17); v1 ^= v2; v2 = emb_rotl(v2, 32); \
    } while (0)

// ASCII lower-casing (the caller lower-cases non-ASCII text with Unicode rules before it gets here)
__device__ __forceinline__ uint32_t emb_lower(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32u : c; }
__device__ __forceinline__ bool emb_is_token(uint32_t c) {  // after lower-casing
    return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_';
}
// Python's str.isspace() over ASCII: \t \n \v \f \r, \x1c-\x1f and the space
__device__ __forceinline__ bool emb_is_space(uint32_t c) { return (c >= 9 && c <= 13) || (c >= 28 && c <= 32); }

struct TextParams {
    const uint8_t* text;       // chunk bytes (device)
    const int64_t* offsets;    // [n + 1] byte offsets into `text`
    uint32_t n;                // records in this chunk
    int skip_blank;
    uint8_t* keep;             // [n] 1 = record is indexed
    uint32_t* block_count;     // [blocks + 1] kept records per CTA block (exclusive-scanned between the kernels)
    // embed
    int dim;                   // logical dimension (= index d)
    int d_pad;                 // stored elements per row
    int store;                 // 0 fp32, 1 bf16
    int normalize;
    uint8_t* rows;             // index row storage (row 0)
    uint64_t pitch_bytes;
    int64_t* ids;              // index id storage (row 0)
    const int64_t* ids_in;     // [n] explicit record ids, or null: first_id + record position
    int64_t first_id;          // id of this chunk's record 0 when ids_in is null
    const unsigned long long* row_base;  // device counter: rows already written (absolute row position)
};

// kernel 1: which records are indexed, and how many per CTA block
__global__ void __launch_bounds__(EMB_THREADS) text_classify_kernel(const TextParams p) {
    __shared__ unsigned int s_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    const uint32_t rec0 = blockIdx.x * EMB_BLOCK_RECS;
    unsigned int mine = 0;
    for (uint32_t r = rec0 + warp; r < rec0 + EMB_BLOCK_RECS && r < p.n; r += EMB_THREADS / 32) {
        bool keep = true;
        if (p.skip_blank) {
            const int64_t b0 = p.offsets[r], b1 = p.offsets[r + 1];
            bool nonspace = false;
            for (int64_t b = b0 + lane; __any_sync(B200_FULL_MASK, b < b1) && !nonspace; b += 32) {
                const bool ns = b < b1 && !emb_is_space(p.text[b]);
                nonspace = __any_sync(B200_FULL_MASK, ns);
            }
            keep = nonspace;
        }
        if (lane == 0) {
            p.keep[r] = keep ? 1 : 0;
            mine += keep ? 1u : 0u;
        }
    }
    if (lane == 0 && mine) atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) p.block_count[blockIdx.x] = s_count;
}

// kernel 3 (after the scan of block_count): embed + normalise + store
template <int STORE>
__global__ void __launch_bounds__(EMB_THREADS) text_embed_kernel(const TextParams p) {
    extern __shared__ __align__(16) float s_vec[];  // [8 warps][dim_s]
    __shared__ uint32_t s_pos[EMB_BLOCK_RECS];
    __shared__ uint32_t s_wtot[EMB_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t rec0 = blockIdx.x * EMB_BLOCK_RECS;
    // exclusive scan of the block's keep flags: position of every kept record among the block's kept records
    {
        const uint32_t r = rec0 + threadIdx.x;
        const uint32_t v = r < p.n ? p.keep[r] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(B200_FULL_MASK, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_wtot[warp] = x;
        __syncthreads();
        uint32_t before = 0;
        for (int w = 0; w < warp; ++w) before += s_wtot[w];
        s_pos[threadIdx.x] = v ? before + x - v : 0xFFFFFFFFu;
    }
    __syncthreads();
    const uint64_t base = *p.row_base + p.block_count[blockIdx.x];
    const int dim_s = (p.d_pad + 3) & ~3;
    float* vec = s_vec + (size_t)warp * dim_s;
    const int nchunk = (p.dim + 3) >> 2;
    for (uint32_t lr = warp; lr < EMB_BLOCK_RECS && rec0 + lr < p.n; lr += EMB_THREADS / 32) {
        const uint32_t pos = s_pos[lr];
        if (pos == 0xFFFFFFFFu) continue;
        const uint32_t r = rec0 + lr;
        for (int c = lane; c < dim_s; c += 32) vec[c] = 0.0f;
        __syncwarp();
        const int64_t b0 = p.offsets[r], b1 = p.offsets[r + 1];
        bool prev_tok = false;  // the byte before this window belonged to a token
        for (int64_t w0 = b0; w0 < b1; w0 += 32) {
            const int64_t b = w0 + lane;
            const uint32_t c = b < b1 ? emb_lower(p.text[b]) : 0u;
            const bool tok = b < b1 && emb_is_token(c);
            const unsigned mask = __ballot_sync(B200_FULL_MASK, tok);
            const unsigned prev = (mask << 1) | (prev_tok ? 1u : 0u);
            const bool start = tok && !((prev >> lane) & 1u);
            prev_tok = (mask >> 31) & 1u;
            if (start) {
                // this lane hashes the token that starts at byte b (it may run past the window)
                uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull, v2 = 0x6c7967656e657261ull, v3 = 0x7465646279746573ull;
                uint64_t m = (uint64_t)c;
                uint32_t len = 1;
                for (int64_t t = b + 1; t < b1; ++t) {
                    const uint32_t ct = emb_lower(p.text[t]);
                    if (!emb_is_token(ct)) break;
                    if ((len & 7u) == 0u) {  // a full 8-byte word precedes this byte
                        v3 ^= m; EMB_SIPROUND; v0 ^= m;
                        m = 0;
                    }
                    m |= (uint64_t)ct << (8 * (len & 7u));
                    ++len;
                }
                if ((len & 7u) == 0u) {  // the token ended exactly on a word boundary: that word is still pending
                    v3 ^= m; EMB_SIPROUND; v0 ^= m;
                    m = 0;
                }
                const uint64_t last = ((uint64_t)len << 56) | m;
                v3 ^= last; EMB_SIPROUND; v0 ^= last;
                v2 ^= 0xff;
                EMB_SIPROUND; EMB_SIPROUND; EMB_SIPROUND;
                long long h = (long long)((v0 ^ v1) ^ (v2 ^ v3));
                if (h == -1) h = -2;  // CPython reserves -1
                const uint64_t a = h < 0 ? (uint64_t)(-(h + 1)) + 1ull : (uint64_t)h;  // abs() without overflow
                atomicAdd(vec + (uint32_t)(a % (uint64_t)p.dim), (h & 1) ? 1.0f : -1.0f);
            }
        }
        __syncwarp();
        // K1: norm with lane l owning the 4-element chunks l, l+32, ... (ascending fmaf, xor butterfly), true division
        float nrm = 1.0f;
        bool zero = false;
        if (p.normalize) {
            float acc = 0.0f;
            for (int ch = lane; ch < nchunk; ch += 32)
                for (int e = 4 * ch; e < 4 * ch + 4 && e < p.dim; ++e) acc = fmaf(vec[e], vec[e], acc);
            acc = warp_sum_xor(acc);
            nrm = __fsqrt_rn(acc);
            zero = ((double)nrm <= 1e-8);
        }
        uint8_t* dst = p.rows + (base + pos) * p.pitch_bytes;
        for (int e = lane; e < p.d_pad; e += 32) {
            float v = e < p.dim ? vec[e] : 0.0f;
            if (p.normalize) v = (zero || e >= p.dim) ? 0.0f : __fdiv_rn(v, nrm);
            if (STORE == 0) reinterpret_cast<float*>(dst)[e] = v;
            else reinterpret_cast<__nv_bfloat16*>(dst)[e] = __float2bfloat16_rn(v);
        }
        if (lane == 0 && p.ids) p.ids[base + pos] = p.ids_in ? p.ids_in[r] : p.first_id + (int64_t)r;
        __syncwarp();
    }
}

// kernel 4: rows written so far += kept records of this chunk (block_count[blocks] after the exclusive scan)
__global__ void text_advance_kernel(unsigned long long* row_base, const uint32_t* block_count, uint32_t blocks) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *row_base += block_count[blocks];
}
