/*
 * flat_oracle.c — CPU oracle for the flat vector recall path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's library; the product (c99_vectordb_b200/) never does.
 *
 * PARITY UNPINNED for the search arithmetic: the reference (memo_cli.py) delegates every vector
 * operation to the third-party dependency faiss-cpu (>=1.13.0, pyproject.toml:11, no exact pin,
 * not vendored, not installed in this image, no wheel available), and ships no tests, golden
 * vectors or fixtures for this path.  The search functions below restate faiss's published
 * flat-index semantics [upstream: faiss/IndexFlat.cpp, utils/distances.cpp, utils/Heap.h,
 * impl/ResultHandler.h, IndexIDMap.cpp] as anchored on the reference's own call sites:
 *
 *   memo_cli.py:131-135  normalize()                     -> oracle_normalize_rows
 *   memo_cli.py:244-248  create_index() (IDMap2 over a flat index, restated IP / L2)
 *   memo_cli.py:282,:437 index.add_with_ids(x, ids)      -> rows + ids arrays passed to search
 *   memo_cli.py:288-298  search_all(): index.search(q,k) -> oracle_search
 *   memo_cli.py:294-297  id < 0 entries dropped          -> -1 padding produced here
 *
 * What IS pinned: oracle_normalize_rows and the adapter-level behaviour are checked against
 * outputs of the reference's own Python functions run in the build container
 * (tests/golden/make_golden.py, fixtures committed under tests/golden/).
 *
 * Semantics restated:
 *   - IP score  = sum x_i*y_i, larger is better, results descending.
 *   - L2 score  = sum (x_i-y_i)^2 (squared), smaller is better, results ascending.
 *   - a row enters the result only on a strict improvement over the current k-th best, scanning
 *     rows in storage order (faiss heap: "if (top < score) replace"), so at the k-th boundary the
 *     earliest row wins; heaps start at -FLT_MAX (IP) / +FLT_MAX (L2), so NaN and -/+inf-worse
 *     scores never enter; unfilled slots are id -1 with that sentinel score.
 *   - stated exact-tie rule of this project: (score best-first, then smaller row position
 *     first).  faiss's final heap_reorder is believed to order exact IP ties by descending id
 *     [upstream, unverifiable here]; that is a documented deviation.
 *   - IndexIDMap2: result labels are id_map[row] for row >= 0.
 *
 * Summation order.  Upstream's fvec_inner_product / fvec_L2sqr are auto-vectorised loops whose
 * order is compiler-defined, so distances are only comparable to ~1e-5 relative.  Two orders are
 * provided: ORDER_SIMD (32 partial sums, multiply then add — a typical AVX2 4x-unrolled build)
 * and ORDER_DEVICE (exactly the B200 kernels' order, scan_topk.cuh, for bit-exact checks).
 * oracle_scores_f64 is the fp64 "truth" used to adjudicate near-ties.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_METRIC_IP 0
#define ORACLE_METRIC_L2 1
#define ORACLE_ORDER_SIMD 0
#define ORACLE_ORDER_DEVICE 1   /* 32 lanes per row */
#define ORACLE_ORDER_DEVICE16 2 /* 16 lanes per row: rows of <= 48 sixteen-byte chunks (scan_topk.cuh LPR) */
#define ORACLE_ORDER_DEVICE8 3  /* 8 lanes per row: rows of 8 or 16 sixteen-byte chunks */

int oracle_version(void) { return 1; }
int oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* Thread count of the OpenMP regions below.  bench.py sets it explicitly: under torchrun the environment carries
 * OMP_NUM_THREADS=1, which would silently turn the "all host cores" baseline into a one-core one. */
void oracle_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ---- synthetic generator: integer-exact twin of csrc/common.cuh b200_synth_value ------------ */
static inline uint32_t synth_bits(uint64_t seed, uint64_t ctr) {
    uint64_t z = ctr + seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return (uint32_t)(z >> 32);
}
void oracle_synth_rows(float* out, int64_t n, int d, uint64_t seed, int64_t first_row) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r)
        for (int c = 0; c < d; ++c) {
            uint32_t b = synth_bits(seed, (uint64_t)(first_row + r) * (uint64_t)d + (uint64_t)c);
            out[r * (int64_t)d + c] = (float)(b >> 8) * (1.0f / 8388608.0f) - 1.0f;
        }
}

/* ---- bf16 storage model: round-to-nearest-even to bf16, widened back to fp32 ---------------- */
static inline float round_bf16(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) { /* NaN: keep quiet NaN */
        u = (u | 0x00400000u) & 0xffff0000u;
    } else {
        uint32_t lsb = (u >> 16) & 1u;
        u += 0x7fffu + lsb;
        u &= 0xffff0000u;
    }
    float y;
    memcpy(&y, &u, 4);
    return y;
}
void oracle_round_bf16(float* x, int64_t count) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < count; ++i) x[i] = round_bf16(x[i]);
}

/* ---- scores ----------------------------------------------------------------------------------- */
static float score_simd(int metric, const float* q, const float* y, int d) {
    float acc[32];
    for (int j = 0; j < 32; ++j) acc[j] = 0.0f;
    int i = 0;
    for (; i + 32 <= d; i += 32)
        for (int j = 0; j < 32; ++j) {
            if (metric == ORACLE_METRIC_IP) {
                acc[j] += q[i + j] * y[i + j];
            } else {
                float t = q[i + j] - y[i + j];
                acc[j] += t * t;
            }
        }
    for (int w = 16; w >= 1; w >>= 1)
        for (int j = 0; j < w; ++j) acc[j] += acc[j + w];
    float s = acc[0];
    for (; i < d; ++i) {
        if (metric == ORACLE_METRIC_IP) {
            s += q[i] * y[i];
        } else {
            float t = q[i] - y[i];
            s += t * t;
        }
    }
    return s;
}

/* The B200 kernels' order: elements are grouped in chunks of `chunk` (4 for fp32 rows, 8 for bf16
 * rows); lane l of `lanes` owns chunks l, l+lanes, ... and accumulates them in ascending element order
 * with a fused multiply-add; lanes are then combined by the xor butterfly lanes/2,...,1.  lanes is 32, or 16
 * for rows of <= 48 chunks (<= 768 bytes), where the kernels put two rows on one warp. */
static float score_device(int metric, const float* q, const float* y, int d, int chunk, int lanes) {
    float lane[32];
    for (int l = 0; l < 32; ++l) lane[l] = 0.0f;
    int nchunk = (d + chunk - 1) / chunk;
    for (int c = 0; c < nchunk; ++c) {
        int l = c % lanes;
        float a = lane[l];
        for (int e = c * chunk; e < (c + 1) * chunk && e < d; ++e) {
            if (metric == ORACLE_METRIC_IP) {
                a = fmaf(y[e], q[e], a);
            } else {
                float t = y[e] - q[e];
                a = fmaf(t, t, a);
            }
        }
        lane[l] = a;
    }
    /* padding elements (d not a multiple of chunk) contribute fmaf(0,0,a) = a (IP) or
     * fmaf(0-0,0-0,a) = a (L2): no effect, skipped above. */
    for (int m = lanes / 2; m >= 1; m >>= 1) {
        float t[32];
        for (int l = 0; l < lanes; ++l) t[l] = lane[l] + lane[l ^ m];
        for (int l = 0; l < lanes; ++l) lane[l] = t[l];
    }
    return lane[0];
}

static inline float score_one(int metric, int order, int chunk, const float* q, const float* y, int d) {
    if (order == ORACLE_ORDER_DEVICE) return score_device(metric, q, y, d, chunk, 32);
    if (order == ORACLE_ORDER_DEVICE16) return score_device(metric, q, y, d, chunk, 16);
    if (order == ORACLE_ORDER_DEVICE8) return score_device(metric, q, y, d, chunk, 8);
    return score_simd(metric, q, y, d);
}

void oracle_scores(int metric, int order, int chunk, const float* db, int64_t n, int d, const float* q,
                   float* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) out[r] = score_one(metric, order, chunk, q, db + r * (int64_t)d, d);
}

void oracle_scores_f64(int metric, const float* db, int64_t n, int d, const float* q, double* out) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        const float* y = db + r * (int64_t)d;
        double s = 0.0;
        for (int i = 0; i < d; ++i) {
            if (metric == ORACLE_METRIC_IP) {
                s += (double)q[i] * (double)y[i];
            } else {
                double t = (double)q[i] - (double)y[i];
                s += t * t;
            }
        }
        out[r] = s;
    }
}

/* ---- top-k ------------------------------------------------------------------------------------ */
typedef struct {
    float s;
    int64_t row;
} cand_t;

/* a is strictly better than b under (score best-first, smaller row first) */
static inline int better(int metric, float sa, int64_t ra, float sb, int64_t rb) {
    if (metric == ORACLE_METRIC_IP) {
        if (sa > sb) return 1;
        if (sa < sb) return 0;
    } else {
        if (sa < sb) return 1;
        if (sa > sb) return 0;
    }
    return ra < rb;
}
static inline int score_valid(int metric, float s) {
    return metric == ORACLE_METRIC_IP ? (s > -FLT_MAX) : (s < FLT_MAX);
}

/* binary heap with the WORST kept candidate at the root */
static void heap_sift_down(int metric, cand_t* h, int64_t n, int64_t i) {
    for (;;) {
        int64_t l = 2 * i + 1, r = l + 1, w = i;
        if (l < n && better(metric, h[w].s, h[w].row, h[l].s, h[l].row)) w = l;
        if (r < n && better(metric, h[w].s, h[w].row, h[r].s, h[r].row)) w = r;
        if (w == i) return;
        cand_t t = h[i];
        h[i] = h[w];
        h[w] = t;
        i = w;
    }
}
static void heap_sift_up(int metric, cand_t* h, int64_t i) {
    while (i > 0) {
        int64_t p = (i - 1) / 2;
        if (better(metric, h[p].s, h[p].row, h[i].s, h[i].row)) {
            cand_t t = h[i];
            h[i] = h[p];
            h[p] = t;
            i = p;
        } else
            return;
    }
}
typedef struct {
    cand_t* h;
    int64_t size, k;
} topk_t;
static inline void topk_push(int metric, topk_t* t, float s, int64_t row) {
    if (!score_valid(metric, s)) return;
    if (t->size < t->k) {
        t->h[t->size].s = s;
        t->h[t->size].row = row;
        heap_sift_up(metric, t->h, t->size);
        t->size++;
    } else if (better(metric, s, row, t->h[0].s, t->h[0].row)) {
        t->h[0].s = s;
        t->h[0].row = row;
        heap_sift_down(metric, t->h, t->size, 0);
    }
}
/* heap -> best-first order in place (repeatedly pop the worst to the back) */
static void topk_sort(int metric, topk_t* t) {
    for (int64_t n = t->size; n > 1; --n) {
        cand_t w = t->h[0];
        t->h[0] = t->h[n - 1];
        t->h[n - 1] = w;
        heap_sift_down(metric, t->h, n - 1, 0);
    }
}
static void topk_emit(int metric, topk_t* t, const int64_t* ids, float* D, int64_t* I) {
    topk_sort(metric, t);
    for (int64_t i = 0; i < t->k; ++i) {
        if (i < t->size) {
            D[i] = t->h[i].s;
            I[i] = ids ? ids[t->h[i].row] : t->h[i].row;
        } else {
            D[i] = metric == ORACLE_METRIC_IP ? -FLT_MAX : FLT_MAX;
            I[i] = -1;
        }
    }
}

/* index.search(q[nq,d], k) over rows db[n,d] with optional id map.  OpenMP over QUERIES only, as
 * faiss's sequential path does (nq < 20) [upstream] — a single query runs on one core.
 * Returns 0, or 1 on allocation failure. */
int oracle_search(int metric, int order, int chunk, const float* db, int64_t n, int d, const int64_t* ids,
                  const float* q, int64_t nq, int64_t k, float* D, int64_t* I) {
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int64_t qi = 0; qi < nq; ++qi) {
        int64_t cap = k < n ? k : n; /* the heap never needs more than min(k, n) slots */
        if (cap < 1) cap = 1;
        topk_t t;
        t.h = (cand_t*)malloc((size_t)cap * sizeof(cand_t));
        t.size = 0;
        t.k = cap;
        if (!t.h) {
            err = 1;
            continue;
        }
        const float* qv = q + qi * (int64_t)d;
        for (int64_t r = 0; r < n; ++r)
            topk_push(metric, &t, score_one(metric, order, chunk, qv, db + r * (int64_t)d, d), r);
        t.k = k; /* emit pads positions size..k-1 */
        topk_emit(metric, &t, ids, D + qi * k, I + qi * k);
        free(t.h);
    }
    return err;
}

/* "best CPU" variant for the reported baseline: rows split across all threads for one query at a
 * time (not how faiss runs nq = 1; our extension, labelled as such in bench.py). */
int oracle_search_rowpar(int metric, int order, int chunk, const float* db, int64_t n, int d,
                         const int64_t* ids, const float* q, int64_t nq, int64_t k, float* D, int64_t* I) {
    int T = oracle_max_threads();
    int64_t cap = k < n ? k : n;
    if (cap < 1) cap = 1;
    cand_t* all = (cand_t*)malloc((size_t)T * (size_t)cap * sizeof(cand_t));
    int64_t* sizes = (int64_t*)malloc((size_t)T * sizeof(int64_t));
    cand_t* fin = (cand_t*)malloc((size_t)cap * sizeof(cand_t));
    if (!all || !sizes || !fin) {
        free(all);
        free(sizes);
        free(fin);
        return 1;
    }
    for (int64_t qi = 0; qi < nq; ++qi) {
        const float* qv = q + qi * (int64_t)d;
#pragma omp parallel num_threads(T)
        {
#ifdef _OPENMP
            int tid = omp_get_thread_num(), nt = omp_get_num_threads();
#else
            int tid = 0, nt = 1;
#endif
            int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
            topk_t t;
            t.h = all + (size_t)tid * cap;
            t.size = 0;
            t.k = cap;
            for (int64_t r = lo; r < hi; ++r)
                topk_push(metric, &t, score_one(metric, order, chunk, qv, db + r * (int64_t)d, d), r);
            sizes[tid] = t.size;
#pragma omp single
            for (int i = nt; i < T; ++i) sizes[i] = 0;
        }
        topk_t f;
        f.h = fin;
        f.size = 0;
        f.k = cap;
        for (int tI = 0; tI < T; ++tI)
            for (int64_t j = 0; j < sizes[tI]; ++j) {
                cand_t c = all[(size_t)tI * cap + j];
                topk_push(metric, &f, c.s, c.row);
            }
        topk_t out = f;
        out.k = k;
        topk_emit(metric, &out, ids, D + qi * k, I + qi * k);
    }
    free(all);
    free(sizes);
    free(fin);
    return 0;
}

/* ---- normalize (memo_cli.py:131-135) --------------------------------------------------------- */
/* n = np.linalg.norm(v) = sqrt(dot(v,v)) in fp32 (numpy routes 1-D float32 dot through BLAS sdot,
 * whose summation order is implementation-defined; ORDER_SIMD models it, ORDER_DEVICE is the K1
 * kernel's order); n <= 1e-8 (a DOUBLE comparison) -> zeros; else v / n (true division). */
void oracle_normalize_rows(float* x, int64_t n, int d, int order) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        float* v = x + r * (int64_t)d;
        float ss = order == ORACLE_ORDER_DEVICE ? score_device(ORACLE_METRIC_IP, v, v, d, 4, 32)
                                                : score_simd(ORACLE_METRIC_IP, v, v, d);
        float nrm = sqrtf(ss);
        if ((double)nrm <= 1e-8) {
            for (int i = 0; i < d; ++i) v[i] = 0.0f;
        } else {
            for (int i = 0; i < d; ++i) v[i] = v[i] / nrm;
        }
    }
}

/* ---- shard merge (K4 restatement) ------------------------------------------------------------- */
/* G best-first lists per query, shard-major [G,nq,k]; ties: lower shard, then earlier position. */
void oracle_merge_topk(int metric, int G, int64_t nq, int64_t k, const float* Dp, const int64_t* Ip, float* Do,
                       int64_t* Io) {
    for (int64_t q = 0; q < nq; ++q) {
        int64_t* pos = (int64_t*)calloc((size_t)G, sizeof(int64_t));
        for (int64_t o = 0; o < k; ++o) {
            int best = -1;
            for (int g = 0; g < G; ++g) {
                if (pos[g] >= k) continue;
                float sg0 = Dp[((int64_t)g * nq + q) * k + pos[g]];
                if (!score_valid(metric, sg0)) continue; /* padding (sentinel score): this shard is exhausted; negative ids are legal */
                if (best < 0) {
                    best = g;
                    continue;
                }
                float sg = Dp[((int64_t)g * nq + q) * k + pos[g]];
                float sb = Dp[((int64_t)best * nq + q) * k + pos[best]];
                int strictly = metric == ORACLE_METRIC_IP ? (sg > sb) : (sg < sb);
                if (strictly) best = g;
            }
            if (best < 0) {
                Do[q * k + o] = metric == ORACLE_METRIC_IP ? -FLT_MAX : FLT_MAX;
                Io[q * k + o] = -1;
            } else {
                Do[q * k + o] = Dp[((int64_t)best * nq + q) * k + pos[best]];
                Io[q * k + o] = Ip[((int64_t)best * nq + q) * k + pos[best]];
                pos[best]++;
            }
        }
        free(pos);
    }
}
