// embed.cu — host-side bulk hashing-trick embedder.
//
// Replaces the per-record Python loop of embed_text_hash (memo_cli.py:158-166): tokens
// [a-z0-9_]+ of the lower-cased text, h = hash(token), vec[abs(h) % dim] += (h & 1) ? +1 : -1.
// The reference uses Python's builtin hash(), which is SipHash-1-3 of the token bytes keyed by a
// per-process random secret (so the same text embeds differently in every process — SURVEY.md §0.4).
// This implementation fixes the key to zero, which is exactly what PYTHONHASHSEED=0 selects in
// CPython >= 3.11: vectors are reproducible across processes and bit-identical to the reference run
// under PYTHONHASHSEED=0 (tests/golden/embed.npz).  Pure host code; no device work.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

#include "../../include/b200_flat.h"

static inline uint64_t rotl64(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
#define SIPROUND                                                    \
    do {                                                            \
        v0 += v1; v1 = rotl64(v1, 13); v1 ^= v0; v0 = rotl64(v0, 32); \
        v2 += v3; v3 = rotl64(v3, 16); v3 ^= v2;                    \
        v0 += v3; v3 = rotl64(v3, 21); v3 ^= v0;                    \
        v2 += v1; v1 = rotl64(v1, 17); v1 ^= v2; v2 = rotl64(v2, 32); \
    } while (0)

// SipHash-1-3 with k0 = k1 = 0, then CPython's conversion to Py_hash_t (-1 is reserved -> -2)
static int64_t py_hash_seed0(const uint8_t* in, size_t len) {
    if (len == 0) return 0;
    uint64_t v0 = 0x736f6d6570736575ull, v1 = 0x646f72616e646f6dull, v2 = 0x6c7967656e657261ull, v3 = 0x7465646279746573ull;
    uint64_t b = (uint64_t)len << 56;
    size_t n = len;
    while (n >= 8) {
        uint64_t mi;
        memcpy(&mi, in, 8);  // little-endian host
        v3 ^= mi;
        SIPROUND;
        v0 ^= mi;
        in += 8;
        n -= 8;
    }
    uint64_t t = 0;
    for (size_t i = 0; i < n; ++i) t |= (uint64_t)in[i] << (8 * i);
    b |= t;
    v3 ^= b;
    SIPROUND;
    v0 ^= b;
    v2 ^= 0xff;
    SIPROUND;
    SIPROUND;
    SIPROUND;
    int64_t h = (int64_t)((v0 ^ v1) ^ (v2 ^ v3));
    return h == -1 ? -2 : h;
}

static inline bool is_token_byte(uint8_t c) { return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_'; }

extern "C" int64_t b200_py_hash_seed0(const char* bytes, int64_t len) {
    return py_hash_seed0(reinterpret_cast<const uint8_t*>(bytes), (size_t)len);
}

// texts: n lower-cased UTF-8 strings concatenated in `utf8`, text i = [offsets[i], offsets[i+1]).
// out: [n, dim] float32, un-normalised signed bucket counts (K1 normalises them at add time).
static void hash_embed_range(const uint8_t* buf, const int64_t* offsets, int64_t i0, int64_t i1, int dim, float* out) {
    for (int64_t i = i0; i < i1; ++i) {
        float* v = out + (size_t)i * dim;
        for (int c = 0; c < dim; ++c) v[c] = 0.0f;
        int64_t p = offsets[i];
        const int64_t e = offsets[i + 1];
        while (p < e) {
            while (p < e && !is_token_byte(buf[p])) ++p;
            int64_t s = p;
            while (p < e && is_token_byte(buf[p])) ++p;
            if (p > s) {
                int64_t h = py_hash_seed0(buf + s, (size_t)(p - s));
                uint64_t a = h < 0 ? (uint64_t)(-(h + 1)) + 1u : (uint64_t)h;  // abs() without overflow
                v[a % (uint64_t)dim] += (h & 1) ? 1.0f : -1.0f;
            }
        }
    }
}

extern "C" int b200_hash_embed(const char* utf8, const int64_t* offsets, int64_t n, int dim, float* out) {
    if (!utf8 || !offsets || !out || n < 0 || dim <= 0) return 1;
    const uint8_t* buf = reinterpret_cast<const uint8_t*>(utf8);
    // records are independent: split them over the host cores (a rebuild embeds every record, memo_cli.py:276-281)
    unsigned hc = std::thread::hardware_concurrency();
    int threads = (int)std::min<unsigned>(32, std::max<unsigned>(1, hc));
    if (const char* env = getenv("B200_EMBED_THREADS")) {
        int t = atoi(env);
        if (t >= 1 && t <= 256) threads = t;
    }
    if (n < 2048 || threads == 1) {
        hash_embed_range(buf, offsets, 0, n, dim, out);
        return 0;
    }
    const int64_t bytes = offsets[n] - offsets[0];
    std::vector<std::thread> th;
    int64_t i0 = 0;
    for (int t = 0; t < threads && i0 < n; ++t) {  // equal BYTES per thread, not equal record counts
        const int64_t target = offsets[0] + bytes * (t + 1) / threads;
        int64_t i1 = t + 1 == threads ? n : (int64_t)(std::upper_bound(offsets + i0, offsets + n, target) - offsets);
        i1 = std::min<int64_t>(std::max<int64_t>(i1, i0), n);
        if (i1 > i0) th.emplace_back(hash_embed_range, buf, offsets, i0, i1, dim, out);
        i0 = i1;
    }
    if (i0 < n) hash_embed_range(buf, offsets, i0, n, dim, out);
    for (auto& t : th) t.join();
    return 0;
}
