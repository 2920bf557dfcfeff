// cabi.cu — the C ABI of include/b200_flat.h over the sm_100a kernels.
// Single translation unit: nvcc -shared -gencode arch=compute_100a,code=sm_100a (see build.py).
#include "../../include/b200_flat.h"

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <chrono>
#include <vector>

#include "common.cuh"
#include "embed_dev.cuh"
#include "fullrank.cuh"
#include "gemm_topk.cuh"
#include "merge.cuh"
#include "rows.cuh"
#include "scan_topk.cuh"

// one (parity, sender) slot of the fused exchange: 16 spare bytes + up to 8 queries x 256 results of three
// self-validating 8-byte words (id low | epoch, id high | epoch, score bits | epoch)
#define XCHG_SLOT_ENTRIES (8 * B200_FUSED_K_MAX)
#define XCHG_SLOT_BYTES ((size_t)16 + (size_t)XCHG_SLOT_ENTRIES * 24)

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)
#define CKI(call)                \
    do {                         \
        int r_ = (call);         \
        if (r_ != 0) return r_;  \
    } while (0)

// ---------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------
struct b200_index {
    int d = 0, d_pad = 0, metric = 0, store = 0, device = 0;
    size_t pitch = 0;  // bytes per stored row
    int lpr = 32;      // lanes sharing a row in the scan arithmetic: 8 / 16 for short rows (see b200_index_create)
    uint8_t* rows = nullptr;
    int64_t ntotal = 0, capacity = 0;
    int64_t* ids = nullptr;
    int64_t ids_capacity = 0;
    int ids_state = 0;  // 0 undecided, 1 explicit ids, 2 ids == row positions
    cudaStream_t stream = nullptr;
    int num_sms = 0;
    size_t smem_optin = 0, smem_per_sm = 0;
    // scratch
    // per-launch control words and survivor lists of the scan kernel, double-buffered by launch parity so that
    // back-to-back launches may overlap (programmatic dependent launch)
    uint64_t* partials = nullptr;   // [2][grid, QB, k] per-CTA best keys
    size_t partials_cap = 0;        // keys per parity set
    unsigned long long* ctl = nullptr;   // [2][16]: word 0 = {ticket, tile counter}
    uint64_t launch_seq = 0;
    unsigned long long* stamps = nullptr;  // [grid, 8] phase stamps of the last scan launch (option scan_phase_stamps)
    size_t stamps_cap = 0;
    int stamps_grid = 0;
    uint8_t* txt_dev[2] = {nullptr, nullptr};  // device staging of text chunks (K6)
    size_t txt_cap = 0;
    uint8_t* txt_keep = nullptr;
    uint32_t* txt_blocks = nullptr;
    unsigned long long* txt_rowbase = nullptr;
    size_t txt_rec_cap = 0;
    cudaStream_t last_stream = nullptr;  // the stream the handle's scratch was last used on
    cudaEvent_t order_ev = nullptr;
    float* q_dev = nullptr;   // staged host queries
    float* qn_dev = nullptr;  // normalised queries
    size_t q_cap = 0, qn_cap = 0;  // floats
    float* D_dev = nullptr;
    int64_t* I_dev = nullptr;
    size_t out_cap = 0;  // entries
    void* pin = nullptr;  // pinned staging for small searches
    size_t pin_cap = 0;
    float* stage = nullptr;  // device staging for ingest
    size_t stage_cap = 0;    // bytes
    uint32_t* fr_hi = nullptr;  // full-rank: [QB, n] hi keys
    uint32_t* fr_buf[4] = {nullptr, nullptr, nullptr, nullptr};  // keys A/B, rows A/B
    uint32_t* fr_hist = nullptr;
    size_t fr_cap = 0, fr_hi_cap = 0, fr_hist_cap = 0;
    // options
    int64_t opt_variant = B200_SCAN_AUTO, opt_warps = 16, opt_stages = 0, opt_tile_rows = 0,
            opt_ctas_per_sm = 0, opt_evict_first = 0, opt_fullrank_min_k = B200_FUSED_K_MAX + 1,
            opt_normalize_queries = 0, opt_staged_results = 1, opt_qb = 0, opt_dynamic = -1, opt_claim_chunk = 0, opt_fused_tail = -1,
            opt_claim_min = 0, opt_claim_first = 0, opt_pdl = 1, opt_queries_stable = 0, opt_phase_stamps = 0, opt_fuse_query_norm = 1;
    bool cur_norm_q = false;  // the scan launches of the search in flight normalise their queries themselves
    int64_t opt_gemm_min_rows = 4096, opt_gemm_min_nq = 2, opt_gemm_emit_factor = 8, opt_gemm_chunk_tiles = 0, opt_gemm_sample_tiles = 1024, opt_gemm_cta_group = 2,
            opt_gemm_shadow_max_rows = 0,  // > 0: keep at most this many rows of the bf16 shadow resident (streamed beyond; tests)
            opt_prefilter = 0,             // 1: single queries rank the bf16 shadow first (half the bytes), then re-rank exactly
            opt_direct_results = 1,        // host API: small results are written straight into pinned host memory by the kernels
            opt_gemm_rows_form = 1;        // small batches: rows as the M operand, queries resident in shared memory (0: always the 256 x 256 form)
    // read-only statistics of the last batched (K3) search
    int64_t stat_gemm_used = 0, stat_gemm_fallbacks = 0, stat_gemm_cand_total = 0, stat_gemm_pass1_us = 0,
            stat_gemm_pass2_us = 0, stat_gemm_rerank_us = 0, stat_gemm_scan_fallbacks = 0, stat_gemm_pre_us = 0,
            stat_gemm_host_us = 0, stat_gemm_streamed = 0, stat_prefilter_used = 0, stat_prefilter_fallbacks = 0,
            stat_gemm_rows_form = 0;
    // K3 state
    __nv_bfloat16* sh_rows = nullptr;  // bf16 shadow of the rows [ntotal, kpad]
    float* sh_norm2 = nullptr;
    unsigned int* sh_maxnorm = nullptr;
    int64_t sh_valid_rows = -1;        // rows covered by the shadow (-1 = none)
    int64_t sh_failed_rows = -1;       // ntotal at which the shadow allocation last failed (no retry until it changes)
    size_t sh_cap_rows = 0;
    uint8_t* pf_buf = nullptr;         // scratch of the single-query pre-filter (lists, thresholds, extended query)
    size_t pf_cap = 0;
    int* pf_cert_host = nullptr;       // pinned: the certificate of the last pre-filtered search
    bool sh_streamed = false;          // the shadow does not fit: sh_rows is a scratch of sh_cap_rows rows, refilled chunk by chunk per search
    cudaStream_t sh_stream2 = nullptr; // streamed shadow: the converter of chunk c+1 runs here while the GEMM sweeps chunk c
    cudaEvent_t sh_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [0,1] chunk converted, [2,3] chunk swept, [4] fork
    __nv_bfloat16* g_qb = nullptr;     // bf16 queries [m_tiles*128, kpad]
    float* g_qnorm2 = nullptr;
    float* g_tilemax = nullptr;
    float* g_theta = nullptr;
    unsigned int* g_count = nullptr;
    uint32_t* g_cand = nullptr;
    int* g_cert = nullptr;
    size_t g_qb_cap = 0, g_q_cap = 0, g_tilemax_cap = 0, g_cand_cap = 0;
    cudaEvent_t g_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // pass brackets of the last batched search
    bool g_ev_pending = false;  // the brackets of an un-synchronised (row-sharded) batch have not been read yet
    uint8_t* up_pin[2] = {nullptr, nullptr};  // pinned upload ring for bulk host adds
    cudaEvent_t up_ev[2] = {nullptr, nullptr};
    size_t up_chunk = 0;
    // fused multi-GPU exchange
    uint8_t** xchg_peers_dev = nullptr;  // device array of world peer buffer pointers
    int xchg_world = 0, xchg_rank = 0;
    uint32_t xchg_epoch = 0;
    int* xchg_status = nullptr;             // device view of the host-mapped flag below
    volatile int* xchg_status_host = nullptr;  // set by a kernel whose peers never arrived
    bool xchg_active = false;            // the search in flight exchanges
    const uint32_t* cur_mask = nullptr;  // row bitmap of the search in flight (device), or null
    uint32_t* mask_dev = nullptr;        // staging for host masks
    size_t mask_cap = 0;
    int64_t* allow_dev = nullptr;        // allowed-id list of a search by ids (device copy)
    size_t allow_cap = 0;
    uint32_t* idbits_dev = nullptr;      // bitmap over the id range
    size_t idbits_cap = 0;               // words
    long long* ids_minmax_dev = nullptr; // [min, max] of the id map
    int64_t ids_min = 0, ids_max = -1;
    int64_t ids_minmax_rows = -1;        // ntotal the cached range was computed for (-1 = stale)
    int64_t launches = 0;
};

static int use_device(b200_index* ix) {
    CK(cudaSetDevice(ix->device));
    return 0;
}

// All device work of a handle shares its scratch (control words, survivor lists, staged queries ...), so work enqueued
// on a different stream than the previous call's must be ordered behind it: an event recorded at the tail of the stream
// used last (which still exists: streams handed to this library must outlive the handle's use of them) and waited for
// on the new one.  Nothing is recorded while the caller stays on one stream.
static int order_after_previous_stream(b200_index* ix, cudaStream_t st) {
    if (ix->last_stream != st && ix->last_stream != nullptr) {
        CK(cudaEventRecord(ix->order_ev, ix->last_stream));
        CK(cudaStreamWaitEvent(st, ix->order_ev, 0));
    }
    ix->last_stream = st;
    return 0;
}

// device scratch that must not outlive an early error return
struct DevTmp {
    void* p = nullptr;
    ~DevTmp() {
        if (p) cudaFree(p);
    }
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes); }
};

template <typename T>
static int grow(T** p, size_t* cap, size_t need) {
    if (*cap >= need) return 0;
    if (*p) CK(cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    CK(cudaMalloc((void**)p, need * sizeof(T)));
    *cap = need;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// library
// ---------------------------------------------------------------------------------------------
extern "C" int b200_abi_version(void) { return B200_ABI_VERSION; }
extern "C" const char* b200_last_error(void) { return g_err.c_str(); }
extern "C" int b200_device_count(int* out_count) {
    if (!out_count) return fail("out_count is null");
    CK(cudaGetDeviceCount(out_count));
    return 0;
}

// Lanes of a warp that share one row in the scan arithmetic (scan_topk.cuh LPR), from the row length in 16-byte
// chunks.  Whole groups keep their measured choices (8 lanes for 8 / 16 chunks, 16 lanes up to 48 chunks, else the
// full warp); ragged rows of up to 64 chunks (1 KB) take the lane count that wastes the fewest chunk slots, ties
// going to fewer lanes (d = 48 fp32: 12 of 32 lanes busy with a full warp, 12 of 16 slots with 8 lanes).
// oracle/oracle.py: device_lanes restates this rule.
static int pick_lpr(size_t nvec) {
    if (nvec <= 16 && nvec % 8 == 0) return 8;
    if (nvec <= 48 && nvec % 16 == 0) return 16;
    if (nvec % 32 == 0 || nvec > 64) return 32;
    int best = 32;
    size_t best_slots = (nvec + 31) / 32 * 32;
    const size_t s16 = (nvec + 15) / 16 * 16, s8 = (nvec + 7) / 8 * 8;
    if (s16 <= best_slots) best = 16, best_slots = s16;
    if (s8 <= best_slots) best = 8, best_slots = s8;
    return best;
}

extern "C" int b200_index_create(b200_index** out, int d, int metric, int store, int device) {
    if (!out) return fail("out is null");
    *out = nullptr;
    if (d <= 0) return fail("d must be positive, got %d", d);
    if (metric != B200_METRIC_IP && metric != B200_METRIC_L2) return fail("unknown metric %d", metric);
    if (store != B200_STORE_F32 && store != B200_STORE_BF16) return fail("unknown store %d", store);
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (count <= 0) return fail("no CUDA device: this library has no CPU fallback");
    if (device < 0 || device >= count) return fail("device %d out of range (have %d)", device, count);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail("device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                    prop.major, prop.minor);
    b200_index* ix = new b200_index();
    ix->d = d;
    ix->metric = metric;
    ix->store = store;
    ix->device = device;
    ix->d_pad = store == B200_STORE_F32 ? (d + 3) / 4 * 4 : (d + 7) / 8 * 8;
    ix->pitch = (size_t)ix->d_pad * (store == B200_STORE_F32 ? 4 : 2);
    ix->lpr = pick_lpr(ix->pitch / 16);  // part of the index's numerics: fixed at creation
    ix->num_sms = prop.multiProcessorCount;
    ix->smem_optin = prop.sharedMemPerBlockOptin;
    ix->smem_per_sm = prop.sharedMemPerMultiprocessor;
    cudaError_t e = cudaStreamCreateWithFlags(&ix->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc((void**)&ix->ctl, 2 * 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ix->ctl, 0, 2 * 16 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ix->order_ev, cudaEventDisableTiming);
    if (e != cudaSuccess) {
        delete ix;
        return fail("index create: %s", cudaGetErrorString(e));
    }
    *out = ix;
    return 0;
}

extern "C" int b200_index_destroy(b200_index* ix) {
    if (!ix) return 0;
    cudaSetDevice(ix->device);
    if (ix->stream) cudaStreamSynchronize(ix->stream);
    cudaFree(ix->rows);
    cudaFree(ix->ids);
    cudaFree(ix->partials);
    cudaFree(ix->ctl);
    cudaFree(ix->stamps);
    for (int i = 0; i < 2; ++i) cudaFree(ix->txt_dev[i]);
    cudaFree(ix->txt_keep);
    cudaFree(ix->txt_blocks);
    cudaFree(ix->txt_rowbase);
    if (ix->order_ev) cudaEventDestroy(ix->order_ev);
    cudaFree(ix->q_dev);
    cudaFree(ix->qn_dev);
    cudaFree(ix->D_dev);
    cudaFree(ix->I_dev);
    cudaFree(ix->stage);
    cudaFree(ix->fr_hi);
    for (int i = 0; i < 4; ++i) cudaFree(ix->fr_buf[i]);
    cudaFree(ix->fr_hist);
    for (int i = 0; i < 2; ++i) {
        if (ix->up_pin[i]) cudaFreeHost(ix->up_pin[i]);
        if (ix->up_ev[i]) cudaEventDestroy(ix->up_ev[i]);
    }
    cudaFree(ix->xchg_peers_dev);
    cudaFree(ix->mask_dev);
    cudaFree(ix->allow_dev);
    cudaFree(ix->idbits_dev);
    cudaFree(ix->ids_minmax_dev);
    cudaFree(ix->sh_rows);
    cudaFree(ix->pf_buf);
    if (ix->pf_cert_host) cudaFreeHost(ix->pf_cert_host);
    if (ix->sh_stream2) cudaStreamDestroy(ix->sh_stream2);
    for (int i = 0; i < 5; ++i)
        if (ix->sh_ev[i]) cudaEventDestroy(ix->sh_ev[i]);
    cudaFree(ix->sh_norm2);
    cudaFree(ix->sh_maxnorm);
    cudaFree(ix->g_qb);
    cudaFree(ix->g_qnorm2);
    cudaFree(ix->g_tilemax);
    cudaFree(ix->g_theta);
    cudaFree(ix->g_count);
    cudaFree(ix->g_cand);
    cudaFree(ix->g_cert);
    for (int i = 0; i < 5; ++i)
        if (ix->g_ev[i]) cudaEventDestroy(ix->g_ev[i]);
    if (ix->xchg_status_host) cudaFreeHost((void*)ix->xchg_status_host);
    if (ix->pin) cudaFreeHost(ix->pin);
    if (ix->stream) cudaStreamDestroy(ix->stream);
    delete ix;
    return 0;
}

extern "C" int b200_index_reset(b200_index* ix) {
    if (!ix) return fail("null index");
    CKI(use_device(ix));
    CKI(order_after_previous_stream(ix, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    ix->ntotal = 0;
    ix->ids_state = 0;
    ix->sh_valid_rows = -1;
    ix->ids_minmax_rows = -1;
    return 0;
}

static int ensure_capacity(b200_index* ix, int64_t need, bool need_ids) {
    if (need > 0xFFFFFFF0ll) return fail("at most 2^32-16 rows per index shard, asked for %lld", (long long)need);
    if (need > ix->capacity) {
        int64_t cap = std::max<int64_t>(need, std::max<int64_t>(1024, ix->capacity + ix->capacity / 2));
        uint8_t* nr = nullptr;
        cudaError_t e = cudaMalloc((void**)&nr, (size_t)cap * ix->pitch);
        if (e != cudaSuccess && cap > need) {  // retry with the exact size
            cudaGetLastError();
            cap = need;
            e = cudaMalloc((void**)&nr, (size_t)cap * ix->pitch);
        }
        if (e != cudaSuccess)
            return fail("cannot allocate %.2f GB of row storage: %s", (double)cap * ix->pitch / 1e9,
                        cudaGetErrorString(e));
        if (ix->ntotal > 0)
            CK(cudaMemcpyAsync(nr, ix->rows, (size_t)ix->ntotal * ix->pitch, cudaMemcpyDeviceToDevice, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        if (ix->rows) CK(cudaFree(ix->rows));
        ix->rows = nr;
        ix->capacity = cap;
    }
    if (need_ids && need > ix->ids_capacity) {
        int64_t cap = std::max<int64_t>(need, ix->capacity);
        int64_t* ni = nullptr;
        CK(cudaMalloc((void**)&ni, (size_t)cap * sizeof(int64_t)));
        if (ix->ntotal > 0 && ix->ids)
            CK(cudaMemcpyAsync(ni, ix->ids, (size_t)ix->ntotal * sizeof(int64_t), cudaMemcpyDeviceToDevice, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        if (ix->ids) CK(cudaFree(ix->ids));
        ix->ids = ni;
        ix->ids_capacity = cap;
    }
    return 0;
}

extern "C" int b200_index_reserve(b200_index* ix, int64_t n_total) {
    if (!ix) return fail("null index");
    CKI(use_device(ix));
    return ensure_capacity(ix, n_total, ix->ids_state == 1);
}

struct OptName {
    const char* name;
    int64_t b200_index::*field;
};
static const OptName kOpts[] = {
    {"scan_variant", &b200_index::opt_variant},
    {"scan_warps", &b200_index::opt_warps},
    {"scan_stages", &b200_index::opt_stages},
    {"scan_tile_rows", &b200_index::opt_tile_rows},
    {"scan_ctas_per_sm", &b200_index::opt_ctas_per_sm},
    {"scan_l2_evict_first", &b200_index::opt_evict_first},
    {"scan_query_block", &b200_index::opt_qb},
    {"scan_dynamic_tiles", &b200_index::opt_dynamic},
    {"scan_claim_chunk", &b200_index::opt_claim_chunk},
    {"scan_fused_tail", &b200_index::opt_fused_tail},
    {"scan_claim_min", &b200_index::opt_claim_min},
    {"scan_claim_first", &b200_index::opt_claim_first},
    {"scan_pdl", &b200_index::opt_pdl},
    {"queries_stable", &b200_index::opt_queries_stable},
    {"scan_phase_stamps", &b200_index::opt_phase_stamps},
    {"fuse_query_normalize", &b200_index::opt_fuse_query_norm},
    {"gemm_min_nq", &b200_index::opt_gemm_min_nq},
    {"gemm_min_rows", &b200_index::opt_gemm_min_rows},
    {"gemm_emit_factor", &b200_index::opt_gemm_emit_factor},
    {"gemm_chunk_tiles", &b200_index::opt_gemm_chunk_tiles},
    {"gemm_sample_tiles", &b200_index::opt_gemm_sample_tiles},
    {"gemm_cta_group", &b200_index::opt_gemm_cta_group},
    {"gemm_shadow_max_rows", &b200_index::opt_gemm_shadow_max_rows},
    {"prefilter", &b200_index::opt_prefilter},
    {"gemm_rows_form", &b200_index::opt_gemm_rows_form},
    {"host_direct_results", &b200_index::opt_direct_results},
    {"stat_gemm_rows_form", &b200_index::stat_gemm_rows_form},
    {"stat_prefilter_used", &b200_index::stat_prefilter_used},
    {"stat_prefilter_fallbacks", &b200_index::stat_prefilter_fallbacks},
    {"stat_gemm_streamed", &b200_index::stat_gemm_streamed},
    {"stat_gemm_used", &b200_index::stat_gemm_used},
    {"stat_gemm_fallbacks", &b200_index::stat_gemm_fallbacks},
    {"stat_gemm_scan_fallbacks", &b200_index::stat_gemm_scan_fallbacks},
    {"stat_gemm_cand_total", &b200_index::stat_gemm_cand_total},
    {"stat_gemm_pass1_us", &b200_index::stat_gemm_pass1_us},
    {"stat_gemm_pass2_us", &b200_index::stat_gemm_pass2_us},
    {"stat_gemm_rerank_us", &b200_index::stat_gemm_rerank_us},
    {"stat_gemm_pre_us", &b200_index::stat_gemm_pre_us},
    {"stat_gemm_host_us", &b200_index::stat_gemm_host_us},
    {"fullrank_min_k", &b200_index::opt_fullrank_min_k},
    {"normalize_queries", &b200_index::opt_normalize_queries},
    {"host_staged_results", &b200_index::opt_staged_results},
};
extern "C" int b200_index_set_option(b200_index* ix, const char* name, int64_t value) {
    if (!ix || !name) return fail("null argument");
    for (const OptName& o : kOpts)
        if (strcmp(o.name, name) == 0) {
            ix->*(o.field) = value;
            return 0;
        }
    return fail("unknown option '%s'", name);
}
extern "C" int b200_index_get_option(b200_index* ix, const char* name, int64_t* out_value) {
    if (!ix || !name || !out_value) return fail("null argument");
    if (ix->g_ev_pending && strncmp(name, "stat_gemm_", 10) == 0 && ix->g_ev[3] && cudaEventQuery(ix->g_ev[3]) == cudaSuccess) {
        // pass brackets of a row-sharded batch, read once its stream got there
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ix->g_ev[0], ix->g_ev[1]) == cudaSuccess) ix->stat_gemm_pass1_us = (int64_t)(ms * 1e3);
        if (cudaEventElapsedTime(&ms, ix->g_ev[1], ix->g_ev[2]) == cudaSuccess) ix->stat_gemm_pass2_us = (int64_t)(ms * 1e3);
        if (cudaEventElapsedTime(&ms, ix->g_ev[2], ix->g_ev[3]) == cudaSuccess) ix->stat_gemm_rerank_us = (int64_t)(ms * 1e3);
        if (cudaEventElapsedTime(&ms, ix->g_ev[4], ix->g_ev[0]) == cudaSuccess) ix->stat_gemm_pre_us = (int64_t)(ms * 1e3);
        ix->g_ev_pending = false;
    }
    cudaGetLastError();
    for (const OptName& o : kOpts)
        if (strcmp(o.name, name) == 0) {
            *out_value = ix->*(o.field);
            return 0;
        }
    return fail("unknown option '%s'", name);
}

// ---------------------------------------------------------------------------------------------
// add
// ---------------------------------------------------------------------------------------------
typedef void (*IngestFn)(const IngestParams);
// regs: 0 = general two-pass form; 4 / 8 = whole row held in registers (d <= 512 / d <= 1024)
static IngestFn pick_ingest(int store, int normalize, int vec, int regs) {
#define ING(S, N, V, R) \
    if (store == S && normalize == N && vec == V && regs == R) return ingest_rows_kernel<S, N, V, R>;
#define ING_SN(S, N) ING(S, N, 0, 0) ING(S, N, 1, 0) ING(S, N, 1, 4) ING(S, N, 1, 8)
    ING_SN(0, 0) ING_SN(0, 1) ING_SN(1, 0) ING_SN(1, 1)
#undef ING_SN
#undef ING
    return nullptr;
}

static int ingest_dev(int d, int d_pad, int store, const float* src_dev, uint8_t* dst, size_t dst_pitch,
                      int64_t n, int normalize, int num_sms, cudaStream_t st, int64_t* launches) {
    if (n <= 0) return 0;
    IngestParams p;
    p.src = src_dev;
    p.dst = dst;
    p.pitch_bytes = dst_pitch;
    p.n = (uint64_t)n;
    p.d = d;
    p.d_pad = d_pad;
    int vec = (d % 4 == 0) && (((uintptr_t)src_dev & 15) == 0);
    int regs = 0;
    if (vec && (dst_pitch & 15) == 0 && (((uintptr_t)dst) & 15) == 0 && (store == B200_STORE_BF16 ? d_pad % 8 == 0 : true))
        regs = d <= 512 ? 4 : (d <= 1024 ? 8 : 0);
    IngestFn fn = pick_ingest(store, normalize ? 1 : 0, vec, regs);
    int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)num_sms * 8);
    fn<<<(unsigned)std::max<int64_t>(1, blocks), 256, 0, st>>>(p);
    if (launches) ++*launches;
    CK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------------------------------------
// bulk host -> device upload: pageable user memory is copied by a few host threads into a
// double-buffered PINNED ring, each buffer then moves by one asynchronous DMA — the host copy of
// chunk c+1 overlaps the DMA of chunk c (a plain cudaMemcpy from pageable memory measured 11 GB/s).
// `consume(dev_ptr_of_chunk, chunk_offset_bytes, chunk_bytes)` is called after each chunk's DMA has
// been enqueued (used to launch the K1 ingest kernel on staged chunks).
// ---------------------------------------------------------------------------------------------
static void parallel_memcpy(uint8_t* dst, const uint8_t* src, size_t bytes, int threads) {
    if (threads <= 1 || bytes < ((size_t)4 << 20)) {
        memcpy(dst, src, bytes);
        return;
    }
    std::vector<std::thread> th;
    size_t per = (bytes + threads - 1) / threads;
    per = (per + 4095) & ~(size_t)4095;
    for (int t = 0; t < threads; ++t) {
        size_t off = (size_t)t * per;
        if (off >= bytes) break;
        size_t len = std::min(per, bytes - off);
        th.emplace_back([=] { memcpy(dst + off, src + off, len); });
    }
    for (auto& t : th) t.join();
}

// Threaded pread/pwrite of one contiguous file range (the page cache copy is the cost; 16 threads as for memcpy).
static int parallel_file_io(int fd, uint8_t* buf, size_t bytes, int64_t file_off, int threads, bool write) {
    std::atomic<int> bad{0};
    auto work = [&bad, fd, buf, file_off, write](size_t off, size_t len) {
        while (len > 0) {
            ssize_t r = write ? pwrite(fd, buf + off, len, (off_t)(file_off + (int64_t)off))
                              : pread(fd, buf + off, len, (off_t)(file_off + (int64_t)off));
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) {  // error, or end of file before the promised bytes
                bad.store(r < 0 ? errno : -1);
                return;
            }
            off += (size_t)r;
            len -= (size_t)r;
        }
    };
    if (threads <= 1 || bytes < ((size_t)4 << 20)) {
        work(0, bytes);
    } else {
        std::vector<std::thread> th;
        size_t per = (bytes + threads - 1) / threads;
        per = (per + 4095) & ~(size_t)4095;
        for (int t = 0; t < threads; ++t) {
            size_t off = (size_t)t * per;
            if (off >= bytes) break;
            th.emplace_back(work, off, std::min(per, bytes - off));
        }
        for (auto& t : th) t.join();
    }
    int e = bad.load();
    if (e == -1) return fail("read error: file ends before the bytes its header promises");
    if (e != 0) return fail("%s: %s", write ? "pwrite" : "pread", strerror(e));
    return 0;
}

static int staging_threads() {
    unsigned hc = std::thread::hardware_concurrency();
    int threads = (int)std::min<unsigned>(16, std::max<unsigned>(1, hc));  // host-memory bound: 16 threads measured best
    if (const char* env = getenv("B200_UPLOAD_THREADS")) {
        int t = atoi(env);
        if (t >= 1 && t <= 64) threads = t;
    }
    return threads;
}

static int ensure_pinned_ring(b200_index* ix, size_t chunk) {
    if (!ix->up_pin[0] || ix->up_chunk < chunk) {
        for (int i = 0; i < 2; ++i) {
            if (ix->up_pin[i]) CK(cudaFreeHost(ix->up_pin[i]));
            ix->up_pin[i] = nullptr;
            CK(cudaHostAlloc((void**)&ix->up_pin[i], chunk, cudaHostAllocDefault));
            if (!ix->up_ev[i]) CK(cudaEventCreateWithFlags(&ix->up_ev[i], cudaEventDisableTiming));
        }
        ix->up_chunk = chunk;
    }
    return 0;
}

// Where the bytes of an upload come from: pageable host memory, or a byte range of an open file.
struct HostSource {
    const uint8_t* mem = nullptr;
    int fd = -1;
    int64_t file_off = 0;
    int fill(uint8_t* pinned, size_t off, size_t len, int threads) const {
        if (fd >= 0) return parallel_file_io(fd, pinned, len, file_off + (int64_t)off, threads, false);
        parallel_memcpy(pinned, mem + off, len, threads);
        return 0;
    }
};

template <typename Consume>
static int upload_staged(b200_index* ix, const HostSource& src, size_t bytes, uint8_t* dst_dev, bool dst_is_ring,
                         size_t align, Consume consume) {
    const size_t kChunk = (size_t)64 << 20;
    size_t chunk = kChunk / align * align;
    if (chunk == 0) chunk = align;
    CKI(ensure_pinned_ring(ix, chunk));
    if (bytes < 4 * chunk) {  // medium transfers: at least ~4 chunks so the host copy and the DMA overlap
        size_t c = std::max<size_t>((size_t)2 << 20, bytes / 4);
        c = (c + align - 1) / align * align;
        chunk = std::min(chunk, c);
    }
    const int threads = staging_threads();
    cudaStream_t st = ix->stream;
    size_t off = 0;
    int b = 0;
    bool used[2] = {true, true};  // an earlier upload's DMA may still be reading the ring (never-recorded events return at once)
    while (off < bytes) {
        size_t len = std::min(chunk, bytes - off);
        if (used[b]) CK(cudaEventSynchronize(ix->up_ev[b]));  // the DMA that last read this buffer is done
        CKI(src.fill(ix->up_pin[b], off, len, threads));
        uint8_t* target = dst_is_ring ? dst_dev : dst_dev + off;
        CK(cudaMemcpyAsync(target, ix->up_pin[b], len, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ix->up_ev[b], st));
        used[b] = true;
        CKI(consume(target, off, len));
        off += len;
        b ^= 1;
    }
    return 0;
}

// Bulk device -> pageable host memory through the same pinned ring: the DMA of chunk c+1 overlaps the threaded
// copy-out of chunk c (a plain cudaMemcpy into pageable memory is staged by the driver at roughly half the rate).
static int download_staged(b200_index* ix, const uint8_t* src_dev, size_t bytes, uint8_t* dst_host, cudaStream_t st) {
    if (bytes == 0) return 0;
    CKI(ensure_pinned_ring(ix, (size_t)64 << 20));
    const size_t chunk = std::min<size_t>(ix->up_chunk, (size_t)16 << 20);
    const int threads = staging_threads();
    for (int i = 0; i < 2; ++i) CK(cudaEventSynchronize(ix->up_ev[i]));  // earlier DMAs out of the ring are done
    size_t off = 0, pend_off = 0, pend_len = 0;
    int b = 0, pend_b = -1;
    while (off < bytes || pend_b >= 0) {
        size_t len = off < bytes ? std::min(chunk, bytes - off) : 0;
        if (len) {
            CK(cudaMemcpyAsync(ix->up_pin[b], src_dev + off, len, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(ix->up_ev[b], st));
        }
        if (pend_b >= 0) {
            CK(cudaEventSynchronize(ix->up_ev[pend_b]));
            parallel_memcpy(dst_host + pend_off, ix->up_pin[pend_b], pend_len, threads);
        }
        pend_b = len ? b : -1;
        pend_off = off;
        pend_len = len;
        off += len;
        b ^= 1;
    }
    return 0;
}

static int note_ids(b200_index* ix, bool explicit_ids) {
    int want = explicit_ids ? 1 : 2;
    if (ix->ntotal > 0 && ix->ids_state != 0 && ix->ids_state != want)
        return fail("cannot mix add() and add_with_ids() on one index");
    ix->ids_state = want;
    return 0;
}

// x: host or device rows, or null when `file` names the source (rows at file->file_off; ids, if ids_file_off >= 0,
// at that byte offset of the same file).
static int add_common(b200_index* ix, const float* x, bool x_is_dev, int64_t n, const int64_t* ids,
                      int normalize, const HostSource* file = nullptr, int64_t ids_file_off = -1) {
    if (!ix) return fail("null index");
    if (n < 0) return fail("negative n");
    if (n == 0) return 0;
    if (!x && !file) return fail("x is null");
    const bool with_ids = ids != nullptr || (file && ids_file_off >= 0);
    CKI(use_device(ix));
    CKI(order_after_previous_stream(ix, ix->stream));
    CKI(note_ids(ix, with_ids));
    CKI(ensure_capacity(ix, ix->ntotal + n, with_ids));
    cudaStream_t st = ix->stream;
    if (ids)
        CK(cudaMemcpyAsync(ix->ids + ix->ntotal, ids, (size_t)n * sizeof(int64_t),
                           x_is_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    else if (with_ids) {
        HostSource idsrc = *file;
        idsrc.file_off = ids_file_off;
        CKI(upload_staged(ix, idsrc, (size_t)n * sizeof(int64_t), (uint8_t*)(ix->ids + ix->ntotal), false, sizeof(int64_t),
                          [](uint8_t*, size_t, size_t) { return 0; }));
    }
    uint8_t* dst = ix->rows + (size_t)ix->ntotal * ix->pitch;
    const bool plain = (ix->store == B200_STORE_F32) && !normalize && (ix->d == ix->d_pad);
    const size_t row_bytes = (size_t)ix->d * 4;
    const bool big_host = file || (!x_is_dev && (size_t)n * row_bytes >= ((size_t)8 << 20));
    HostSource src;
    if (file) src = *file;
    else src.mem = (const uint8_t*)x;
    if (plain && !file && (x_is_dev || !big_host)) {
        CK(cudaMemcpyAsync(dst, x, (size_t)n * ix->pitch,
                           x_is_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    } else if (plain) {
        // fp32 rows stored verbatim: pinned ring straight into the resident rows
        CKI(upload_staged(ix, src, (size_t)n * row_bytes, dst, false, row_bytes,
                          [](uint8_t*, size_t, size_t) { return 0; }));
    } else if (x_is_dev) {
        CKI(ingest_dev(ix->d, ix->d_pad, ix->store, x, dst, ix->pitch, n, normalize, ix->num_sms, st, &ix->launches));
    } else {
        // host rows that need K1 (normalise / bf16 / padding): pinned ring -> device staging -> K1
        const size_t chunk_bytes = std::max(row_bytes, (((size_t)64 << 20) / row_bytes) * row_bytes);
        if (ix->stage_cap < chunk_bytes) {
            if (ix->stage) CK(cudaFree(ix->stage));
            ix->stage = nullptr;
            ix->stage_cap = 0;
            CK(cudaMalloc((void**)&ix->stage, chunk_bytes));
            ix->stage_cap = chunk_bytes;
        }
        if (big_host) {
            CKI(upload_staged(ix, src, (size_t)n * row_bytes, (uint8_t*)ix->stage, true, row_bytes,
                              [&](uint8_t* dev_chunk, size_t off, size_t len) {
                                  int64_t r0 = (int64_t)(off / row_bytes), nr = (int64_t)(len / row_bytes);
                                  return ingest_dev(ix->d, ix->d_pad, ix->store, (const float*)dev_chunk,
                                                    dst + (size_t)r0 * ix->pitch, ix->pitch, nr, normalize, ix->num_sms, st,
                                                    &ix->launches);
                              }));
        } else {
            CK(cudaMemcpyAsync(ix->stage, x, (size_t)n * row_bytes, cudaMemcpyHostToDevice, st));
            CKI(ingest_dev(ix->d, ix->d_pad, ix->store, ix->stage, dst, ix->pitch, n, normalize, ix->num_sms, st, &ix->launches));
        }
    }
    CK(cudaStreamSynchronize(st));
    ix->ntotal += n;
    ix->sh_valid_rows = -1;
    return 0;
}

extern "C" int b200_index_add(b200_index* ix, const float* x_host, int64_t n, const int64_t* ids_host,
                              int normalize) {
    return add_common(ix, x_host, false, n, ids_host, normalize);
}
extern "C" int b200_index_add_dev(b200_index* ix, const float* x_dev, int64_t n, const int64_t* ids_dev,
                                  int normalize) {
    return add_common(ix, x_dev, true, n, ids_dev, normalize);
}

struct FdGuard {
    int fd = -1;
    ~FdGuard() {
        if (fd >= 0) close(fd);
    }
};

extern "C" int b200_index_add_file(b200_index* ix, const char* path, int64_t rows_offset, int64_t n,
                                   int64_t ids_offset, int normalize) {
    if (!ix || !path) return fail("null argument");
    if (n < 0 || rows_offset < 0) return fail("bad file range");
    if (n == 0) return 0;
    FdGuard g;
    g.fd = open(path, O_RDONLY | O_CLOEXEC);
    if (g.fd < 0) return fail("open %s: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(g.fd, &sb) != 0) return fail("fstat %s: %s", path, strerror(errno));
    const int64_t row_bytes = (int64_t)ix->d * 4;
    if (rows_offset + n * row_bytes > (int64_t)sb.st_size || (ids_offset >= 0 && ids_offset + n * 8 > (int64_t)sb.st_size))
        return fail("read error: %s is shorter than its header promises", path);
    HostSource src;
    src.fd = g.fd;
    src.file_off = rows_offset;
    return add_common(ix, nullptr, false, n, nullptr, normalize, &src, ids_offset);
}

// rows [0, ntotal) as dense fp32 at rows_offset, ids (int64) at ids_offset when >= 0; the file must exist.
// Device -> pinned ring -> threaded pwrite; the D2H of chunk c+1 overlaps the file write of chunk c.
extern "C" int b200_index_write_file(b200_index* ix, const char* path, int64_t rows_offset, int64_t ids_offset) {
    if (!ix || !path) return fail("null argument");
    if (rows_offset < 0) return fail("bad file offset");
    if (ix->ntotal == 0) return 0;
    CKI(use_device(ix));
    FdGuard g;
    g.fd = open(path, O_WRONLY | O_CLOEXEC);
    if (g.fd < 0) return fail("open %s: %s", path, strerror(errno));
    const int threads = staging_threads();
    cudaStream_t st = ix->stream;
    CK(cudaStreamSynchronize(st));
    const size_t row_bytes = (size_t)ix->d * 4;
    if (ix->store != B200_STORE_F32) {
        // bf16 rows widen to fp32 on the host (lossless); not a hot path
        const int64_t step = std::max<int64_t>(1, (int64_t)(((size_t)64 << 20) / row_bytes));
        std::vector<float> tmp((size_t)std::min<int64_t>(step, ix->ntotal) * ix->d);
        for (int64_t r0 = 0; r0 < ix->ntotal; r0 += step) {
            int64_t nr = std::min<int64_t>(step, ix->ntotal - r0);
            CKI(b200_index_get_rows(ix, r0, nr, tmp.data()));
            CKI(parallel_file_io(g.fd, (uint8_t*)tmp.data(), (size_t)nr * row_bytes, rows_offset + r0 * (int64_t)row_bytes, threads, true));
        }
    } else {
        size_t chunk_rows = std::max<size_t>(1, ((size_t)64 << 20) / row_bytes);
        CKI(ensure_pinned_ring(ix, chunk_rows * row_bytes));
        chunk_rows = ix->up_chunk / row_bytes;
        int64_t pending_r0 = -1, pending_n = 0;
        int pending_b = 0, b = 0;
        for (int64_t r0 = 0; r0 < ix->ntotal || pending_r0 >= 0; r0 += (int64_t)chunk_rows) {
            int64_t nr = r0 < ix->ntotal ? std::min<int64_t>((int64_t)chunk_rows, ix->ntotal - r0) : 0;
            if (nr > 0) {
                CK(cudaMemcpy2DAsync(ix->up_pin[b], row_bytes, ix->rows + (size_t)r0 * ix->pitch, ix->pitch, row_bytes,
                                     (size_t)nr, cudaMemcpyDeviceToHost, st));
                CK(cudaEventRecord(ix->up_ev[b], st));
            }
            if (pending_r0 >= 0) {
                CK(cudaEventSynchronize(ix->up_ev[pending_b]));
                CKI(parallel_file_io(g.fd, ix->up_pin[pending_b], (size_t)pending_n * row_bytes,
                                     rows_offset + pending_r0 * (int64_t)row_bytes, threads, true));
            }
            pending_r0 = nr > 0 ? r0 : -1;
            pending_n = nr;
            pending_b = b;
            b ^= 1;
        }
    }
    if (ids_offset >= 0) {
        std::vector<int64_t> ids((size_t)ix->ntotal);
        CKI(b200_index_get_ids(ix, ids.data()));
        CKI(parallel_file_io(g.fd, (uint8_t*)ids.data(), ids.size() * 8, ids_offset, threads, true));
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// whole .memo files: faiss index serialisation [upstream layout, SURVEY.md App. A.5]
// header = int32 d, int64 ntotal, 2 x int64 (unused), uint8 is_trained, int32 metric (+ float32 metric_arg when > 1)
// ---------------------------------------------------------------------------------------------
struct FileGuard {
    FILE* f = nullptr;
    ~FileGuard() {
        if (f) fclose(f);
    }
};

static int rd(FILE* f, void* dst, size_t n) {
    size_t got = fread(dst, 1, n, f);
    if (got != n) return fail("read error: wanted %zu bytes, got %zu", n, got);
    return 0;
}

static int rd_header(FILE* f, int32_t* d, int64_t* ntotal, int32_t* metric) {
    int64_t dummy[2];
    uint8_t trained;
    CKI(rd(f, d, 4));
    CKI(rd(f, ntotal, 8));
    CKI(rd(f, dummy, 16));
    CKI(rd(f, &trained, 1));
    CKI(rd(f, metric, 4));
    if (*metric > 1) {
        float arg;
        CKI(rd(f, &arg, 4));
    }
    if (*d <= 0 || *ntotal < 0) return fail("corrupt index header");
    return 0;
}

// a serialised std::vector: uint64 count, then count elements; skipped, count returned
static int skip_vector(FILE* f, size_t elem_bytes, uint64_t limit, uint64_t* count) {
    uint64_t n = 0;
    CKI(rd(f, &n, 8));
    if (n > limit) return fail("implausible vector length in index file");
    if (fseeko(f, (off_t)(n * elem_bytes), SEEK_CUR) != 0) return fail("seek failed: %s", strerror(errno));
    *count = n;
    return 0;
}

// IndexHNSWFlat: header, then the HNSW graph (assign_probas double[], cum_nneighbor_per_level int32[], levels int32[ntotal],
// offsets size_t[ntotal+1], neighbors int32[], 5 x int32 scalars), then the flat storage index.  The graph is useless for
// an exact flat index and is skipped [upstream write_HNSW layout, unverified here: every size is checked].
static int skip_hnsw_graph(FILE* f, int64_t ntotal) {
    const uint64_t big = (uint64_t)1 << 40;
    uint64_t n = 0;
    CKI(skip_vector(f, 8, 1u << 16, &n));
    CKI(skip_vector(f, 4, 1u << 16, &n));
    CKI(skip_vector(f, 4, big, &n));
    if (n != (uint64_t)ntotal) return fail("HNSW levels do not match ntotal");
    CKI(skip_vector(f, 8, big, &n));
    if (n != (uint64_t)ntotal + 1) return fail("HNSW offsets do not match ntotal");
    CKI(skip_vector(f, 4, big, &n));
    uint8_t scalars[20];
    return rd(f, scalars, sizeof scalars);
}

extern "C" int b200_memo_probe(const char* path, b200_memo_info* out) {
    if (!path || !out) return fail("null argument");
    memset(out, 0, sizeof *out);
    out->ids_offset = -1;
    FileGuard g;
    g.f = fopen(path, "rb");
    if (!g.f) return fail("could not open %s for reading: %s", path, strerror(errno));
    struct stat sb;
    if (fstat(fileno(g.f), &sb) != 0) return fail("fstat %s: %s", path, strerror(errno));
    char cc[5] = {0, 0, 0, 0, 0};
    CKI(rd(g.f, cc, 4));
    int32_t d = 0, metric = 0;
    int64_t ntotal = 0;
    if (memcmp(cc, "IxMp", 4) == 0 || memcmp(cc, "IxM2", 4) == 0) {
        out->kind = cc[3] == '2' ? 2 : 1;
        CKI(rd_header(g.f, &d, &ntotal, &metric));  // the wrapper's own copy; the nested index decides
        CKI(rd(g.f, cc, 4));
    }
    if (memcmp(cc, "IHNf", 4) == 0) {
        CKI(rd_header(g.f, &d, &ntotal, &metric));
        CKI(skip_hnsw_graph(g.f, ntotal));
        out->from_hnsw = 1;
        CKI(rd(g.f, cc, 4));  // the storage index
    }
    if (memcmp(cc, "IxFI", 4) != 0 && memcmp(cc, "IxF2", 4) != 0 && memcmp(cc, "IxFl", 4) != 0) {
        for (int i = 0; i < 4; ++i)
            if (cc[i] < 32 || cc[i] > 126) cc[i] = '?';
        return fail("Index type '%s' not recognized", cc);
    }
    CKI(rd_header(g.f, &d, &ntotal, &metric));
    uint64_t count = 0;
    CKI(rd(g.f, &count, 8));
    if (count != (uint64_t)ntotal * (uint64_t)d) return fail("flat payload size does not match header");
    out->d = d;
    out->metric = metric;
    out->ntotal = ntotal;
    out->rows_offset = (int64_t)ftello(g.f);
    const int64_t rows_end = out->rows_offset + ntotal * (int64_t)d * 4;
    if (rows_end > (int64_t)sb.st_size) return fail("read error: %s is shorter than its header promises", path);
    if (out->kind) {
        if (fseeko(g.f, (off_t)rows_end, SEEK_SET) != 0) return fail("seek failed: %s", strerror(errno));
        uint64_t n_ids = 0;
        CKI(rd(g.f, &n_ids, 8));
        if (n_ids != (uint64_t)ntotal) return fail("id_map size does not match the nested index");
        out->ids_offset = rows_end + 8;
        if (out->ids_offset + ntotal * 8 > (int64_t)sb.st_size) return fail("read error: %s is shorter than its header promises", path);
    }
    return 0;
}

static int wr(FILE* f, const void* src, size_t n) {
    if (fwrite(src, 1, n, f) != n) return fail("write error: %s", strerror(errno));
    return 0;
}
static int wr_header(FILE* f, int32_t d, int64_t ntotal, int32_t metric) {
    const int64_t dummy[2] = {1 << 20, 1 << 20};
    const uint8_t trained = 1;
    CKI(wr(f, &d, 4));
    CKI(wr(f, &ntotal, 8));
    CKI(wr(f, dummy, 16));
    CKI(wr(f, &trained, 1));
    return wr(f, &metric, 4);
}

extern "C" int b200_memo_write_headers(const char* path, const b200_memo_info* in, int64_t* rows_offset, int64_t* ids_offset) {
    if (!path || !in || !rows_offset || !ids_offset) return fail("null argument");
    if (in->d <= 0 || in->ntotal < 0 || in->kind < 0 || in->kind > 2 || (in->metric != B200_METRIC_IP && in->metric != B200_METRIC_L2))
        return fail("bad index description");
    FileGuard g;
    g.f = fopen(path, "wb");
    if (!g.f) return fail("could not open %s for writing: %s", path, strerror(errno));
    if (in->kind) {
        CKI(wr(g.f, in->kind == 2 ? "IxM2" : "IxMp", 4));
        CKI(wr_header(g.f, in->d, in->ntotal, in->metric));
    }
    CKI(wr(g.f, in->metric == B200_METRIC_IP ? "IxFI" : "IxF2", 4));
    CKI(wr_header(g.f, in->d, in->ntotal, in->metric));
    const uint64_t count = (uint64_t)in->ntotal * (uint64_t)in->d;
    CKI(wr(g.f, &count, 8));
    *rows_offset = (int64_t)ftello(g.f);
    *ids_offset = -1;
    if (in->kind) {
        const int64_t rows_end = *rows_offset + in->ntotal * (int64_t)in->d * 4;
        if (fseeko(g.f, (off_t)rows_end, SEEK_SET) != 0) return fail("seek failed: %s", strerror(errno));
        const uint64_t n_ids = (uint64_t)in->ntotal;
        CKI(wr(g.f, &n_ids, 8));
        *ids_offset = rows_end + 8;
    }
    if (fflush(g.f) != 0) return fail("write error: %s", strerror(errno));
    return 0;
}

extern "C" int b200_index_save(b200_index* ix, const char* path, int kind) {
    if (!ix || !path) return fail("null argument");
    b200_memo_info info;
    memset(&info, 0, sizeof info);
    info.kind = kind;
    info.d = ix->d;
    info.metric = ix->metric;
    info.ntotal = ix->ntotal;
    int64_t rows_off = 0, ids_off = -1;
    CKI(b200_memo_write_headers(path, &info, &rows_off, &ids_off));
    return b200_index_write_file(ix, path, rows_off, ids_off);
}

extern "C" int b200_index_load(b200_index** out, const char* path, int store, int device, b200_memo_info* info_or_null) {
    if (!out || !path) return fail("null argument");
    *out = nullptr;
    b200_memo_info info;
    CKI(b200_memo_probe(path, &info));
    if (info_or_null) *info_or_null = info;
    b200_index* ix = nullptr;
    CKI(b200_index_create(&ix, info.d, info.metric, store, device));
    int rc = b200_index_add_file(ix, path, info.rows_offset, info.ntotal, info.ids_offset, 0);
    if (rc) {
        std::string keep = g_err;  // destroy must not lose the message
        b200_index_destroy(ix);
        g_err = keep;
        return rc;
    }
    if (info.kind && info.ntotal == 0) ix->ids_state = 1;  // an empty id-mapped index stays id-mapped
    *out = ix;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K6: bulk rebuild from texts — hashing-trick embedder + K1 on the device (embed_dev.cuh)
// ---------------------------------------------------------------------------------------------
extern "C" int b200_index_add_texts(b200_index* ix, const char* utf8_host, const int64_t* offsets_host, int64_t n,
                                    const int64_t* ids_host, int64_t first_id, int skip_blank, int normalize, int with_ids,
                                    int64_t* n_added) {
    if (!ix) return fail("null index");
    if (n < 0) return fail("negative n");
    if (n_added) *n_added = 0;
    if (n == 0) return 0;
    if (!utf8_host || !offsets_host) return fail("null buffer");
    if (ix->d > EMB_MAX_DIM) return fail("the device embedder supports d <= %d, the index has d = %d", EMB_MAX_DIM, ix->d);
    for (int64_t i = 0; i < n; ++i)
        if (offsets_host[i + 1] < offsets_host[i]) return fail("offsets must be non-decreasing (record %lld)", (long long)i);
    const bool want_ids = with_ids != 0 || ids_host != nullptr;
    CKI(use_device(ix));
    CKI(order_after_previous_stream(ix, ix->stream));
    CKI(note_ids(ix, want_ids));
    CKI(ensure_capacity(ix, ix->ntotal + n, want_ids));  // worst case: nothing is blank
    cudaStream_t st = ix->stream;
    // chunks of whole records: <= 32 MB of text and <= 2^20 records, through the double-buffered pinned ring
    const size_t kTextChunk = (size_t)32 << 20;
    const int64_t kRecChunk = (int64_t)1 << 20;
    size_t max_rec = 0;
    for (int64_t i = 0; i < n; ++i) max_rec = std::max(max_rec, (size_t)(offsets_host[i + 1] - offsets_host[i]));
    const size_t text_cap = std::max(kTextChunk, max_rec) + 64;
    const size_t buf_bytes = text_cap + (size_t)(kRecChunk + 1) * 8 + (size_t)kRecChunk * 8 + 64;
    CKI(ensure_pinned_ring(ix, buf_bytes));
    if (ix->txt_cap < buf_bytes) {
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < 2; ++i) {
            if (ix->txt_dev[i]) CK(cudaFree(ix->txt_dev[i]));
            ix->txt_dev[i] = nullptr;
        }
        ix->txt_cap = 0;
        for (int i = 0; i < 2; ++i) CK(cudaMalloc((void**)&ix->txt_dev[i], buf_bytes));
        ix->txt_cap = buf_bytes;
    }
    const size_t max_blocks = (size_t)((kRecChunk + EMB_BLOCK_RECS - 1) / EMB_BLOCK_RECS);
    if (ix->txt_rec_cap < (size_t)kRecChunk) {
        CK(cudaStreamSynchronize(st));
        cudaFree(ix->txt_keep);
        cudaFree(ix->txt_blocks);
        ix->txt_keep = nullptr;
        ix->txt_blocks = nullptr;
        ix->txt_rec_cap = 0;
        CK(cudaMalloc((void**)&ix->txt_keep, (size_t)kRecChunk));
        CK(cudaMalloc((void**)&ix->txt_blocks, (max_blocks + 1) * sizeof(uint32_t)));
        ix->txt_rec_cap = (size_t)kRecChunk;
    }
    if (!ix->txt_rowbase) CK(cudaMalloc((void**)&ix->txt_rowbase, sizeof(unsigned long long)));
    const unsigned long long start_row = (unsigned long long)ix->ntotal;
    CK(cudaMemcpyAsync(ix->txt_rowbase, &start_row, sizeof start_row, cudaMemcpyHostToDevice, st));
    const int dim_s = (ix->d_pad + 3) & ~3;
    const size_t esmem = (size_t)(EMB_THREADS / 32) * dim_s * sizeof(float);
    CK(cudaFuncSetAttribute(text_embed_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    CK(cudaFuncSetAttribute(text_embed_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)esmem));
    const int threads = staging_threads();
    int b = 0;
    int64_t r0 = 0;
    while (r0 < n) {
        // the records of this chunk
        int64_t r1 = r0;
        const int64_t byte0 = offsets_host[r0];
        while (r1 < n && r1 - r0 < kRecChunk && (size_t)(offsets_host[r1 + 1] - byte0) <= text_cap - 64) ++r1;
        if (r1 == r0) return fail("record %lld does not fit the text staging buffer", (long long)r0);  // cannot happen: text_cap >= max_rec
        const int64_t cnt = r1 - r0;
        const size_t tbytes = (size_t)(offsets_host[r1] - byte0);
        const size_t off_offsets = (tbytes + 15) & ~(size_t)15;
        const size_t off_ids = off_offsets + (size_t)(cnt + 1) * 8;
        const size_t total = off_ids + (ids_host ? (size_t)cnt * 8 : 0);
        CK(cudaEventSynchronize(ix->up_ev[b]));  // the DMA that last read this pinned buffer is done
        uint8_t* pin = ix->up_pin[b];
        parallel_memcpy(pin, reinterpret_cast<const uint8_t*>(utf8_host) + byte0, tbytes, threads);
        int64_t* po = reinterpret_cast<int64_t*>(pin + off_offsets);
        for (int64_t i = 0; i <= cnt; ++i) po[i] = offsets_host[r0 + i] - byte0;
        if (ids_host) memcpy(pin + off_ids, ids_host + r0, (size_t)cnt * 8);
        uint8_t* dev = ix->txt_dev[b];
        CK(cudaMemcpyAsync(dev, pin, total, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(ix->up_ev[b], st));
        TextParams p;
        memset(&p, 0, sizeof p);
        p.text = dev;
        p.offsets = reinterpret_cast<const int64_t*>(dev + off_offsets);
        p.n = (uint32_t)cnt;
        p.skip_blank = skip_blank;
        p.keep = ix->txt_keep;
        p.block_count = ix->txt_blocks;
        p.dim = ix->d;
        p.d_pad = ix->d_pad;
        p.store = ix->store;
        p.normalize = normalize;
        p.rows = ix->rows;
        p.pitch_bytes = ix->pitch;
        p.ids = want_ids ? ix->ids : nullptr;
        p.ids_in = ids_host ? reinterpret_cast<const int64_t*>(dev + off_ids) : nullptr;
        p.first_id = first_id + r0;
        p.row_base = ix->txt_rowbase;
        const uint32_t blocks = (uint32_t)((cnt + EMB_BLOCK_RECS - 1) / EMB_BLOCK_RECS);
        CK(cudaMemsetAsync(ix->txt_blocks + blocks, 0, sizeof(uint32_t), st));  // becomes the chunk total after the scan
        text_classify_kernel<<<blocks, EMB_THREADS, 0, st>>>(p);
        radix_scan_kernel<<<1, 1024, 0, st>>>(ix->txt_blocks, (uint64_t)blocks + 1);
        if (ix->store == B200_STORE_F32) text_embed_kernel<0><<<blocks, EMB_THREADS, esmem, st>>>(p);
        else text_embed_kernel<1><<<blocks, EMB_THREADS, esmem, st>>>(p);
        text_advance_kernel<<<1, 32, 0, st>>>(ix->txt_rowbase, ix->txt_blocks, blocks);
        ix->launches += 4;
        CK(cudaGetLastError());
        r0 = r1;
        b ^= 1;
    }
    unsigned long long end_row = 0;
    CK(cudaMemcpyAsync(&end_row, ix->txt_rowbase, sizeof end_row, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (n_added) *n_added = (int64_t)(end_row - start_row);
    if (end_row != start_row) {
        ix->ntotal = (int64_t)end_row;
        ix->sh_valid_rows = -1;
        ix->ids_minmax_rows = -1;
    } else if (ix->ntotal == 0) {
        ix->ids_state = 0;  // nothing was added: the index stays undecided
    }
    return 0;
}

extern "C" int b200_synth_rows_dev(float* out_dev, int64_t n, int d, uint64_t seed, int64_t first_row,
                                   int normalize, void* stream) {
    if (!out_dev || n < 0 || d <= 0) return fail("bad argument");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    synth_rows_kernel<<<sms * 8, 256, 0, st>>>(out_dev, (uint64_t)n, (uint32_t)d, seed, (uint64_t)first_row);
    CK(cudaGetLastError());
    if (normalize) {
        // in place: same pitch as the dense source when d % 4 == 0
        if (d % 4 != 0) return fail("in-place synthetic normalisation needs d %% 4 == 0");
        CKI(ingest_dev(d, d, B200_STORE_F32, out_dev, (uint8_t*)out_dev, (size_t)d * 4, n, 1, sms, st, nullptr));
    }
    return 0;
}

extern "C" int b200_index_add_synthetic(b200_index* ix, int64_t n, uint64_t seed, int64_t first_row,
                                        int normalize, int with_ids, int64_t first_id) {
    if (!ix) return fail("null index");
    if (n < 0) return fail("negative n");
    if (n == 0) return 0;
    CKI(use_device(ix));
    CKI(order_after_previous_stream(ix, ix->stream));
    CKI(note_ids(ix, with_ids != 0));
    CKI(ensure_capacity(ix, ix->ntotal + n, with_ids != 0));
    cudaStream_t st = ix->stream;
    const size_t row_bytes = (size_t)ix->d * 4;
    size_t chunk_rows = std::max<size_t>(1, ((size_t)256 << 20) / row_bytes);
    chunk_rows = std::min<size_t>(chunk_rows, (size_t)n);
    if (ix->stage_cap < chunk_rows * row_bytes) {
        if (ix->stage) CK(cudaFree(ix->stage));
        ix->stage = nullptr;
        ix->stage_cap = 0;
        CK(cudaMalloc((void**)&ix->stage, chunk_rows * row_bytes));
        ix->stage_cap = chunk_rows * row_bytes;
    }
    uint8_t* dst = ix->rows + (size_t)ix->ntotal * ix->pitch;
    for (int64_t r0 = 0; r0 < n; r0 += (int64_t)chunk_rows) {
        int64_t nr = std::min<int64_t>((int64_t)chunk_rows, n - r0);
        synth_rows_kernel<<<ix->num_sms * 8, 256, 0, st>>>(ix->stage, (uint64_t)nr, (uint32_t)ix->d, seed,
                                                           (uint64_t)(first_row + r0));
        ++ix->launches;
        CK(cudaGetLastError());
        CKI(ingest_dev(ix->d, ix->d_pad, ix->store, ix->stage, dst + (size_t)r0 * ix->pitch, ix->pitch, nr,
                       normalize, ix->num_sms, st, &ix->launches));
    }
    if (with_ids) {
        iota_ids_kernel<<<ix->num_sms * 4, 256, 0, st>>>(ix->ids + ix->ntotal, (uint64_t)n, first_id);
        ++ix->launches;
        CK(cudaGetLastError());
    }
    CK(cudaStreamSynchronize(st));
    ix->ntotal += n;
    ix->sh_valid_rows = -1;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// search
// ---------------------------------------------------------------------------------------------
// What a scan launch reads: the index's own rows by default; the single-query pre-filter scans the bf16 shadow
// through another view (search_prefilter).
struct ScanView {
    const uint8_t* rows;
    size_t pitch;
    int store, metric, d, d_pad, lpr;
    const int64_t* id_map;  // row -> record id, or null (row positions)
    bool normalize_q;       // L2-normalise the queries while staging them
};
static ScanView index_view(const b200_index* ix) {
    return ScanView{ix->rows, ix->pitch, ix->store, ix->metric, ix->d, ix->d_pad, ix->lpr,
                    ix->ids_state == 1 ? ix->ids : nullptr, ix->cur_norm_q};
}

struct ScanPlan {
    int variant = 0, nw = 0, qb = 0;
    uint32_t tile_rows = 0, stages = 0, tile_bytes = 0, scratch_keys = 0;
    size_t smem = 0;
    int grid = 0;
};

typedef void (*ScanFn)(const ScanParams);
#define SCAN_RB 4
// The scan kernel instantiations live in scan_bulk.cu / scan_ldg.cu (separate translation units so
// the build parallelises): 32 lanes per row with query blocks 1/2/4/8, 16 lanes per row (short rows)
// with query blocks 1 and 8.
#define SCAN_DECL(V) \
    ScanFn b200_pick_scan_##V##_m0_s0(int, int); ScanFn b200_pick_scan_##V##_m0_s1(int, int); \
    ScanFn b200_pick_scan_##V##_m1_s0(int, int); ScanFn b200_pick_scan_##V##_m1_s1(int, int);
SCAN_DECL(bulk) SCAN_DECL(ldg)
#undef SCAN_DECL
static ScanFn pick_scan(int metric, int store, int qb, int variant, int lpr) {
    const int ms = metric * 2 + store;
    if (variant == B200_VARIANT_BULK) {
        switch (ms) {
            case 0: return b200_pick_scan_bulk_m0_s0(qb, lpr);
            case 1: return b200_pick_scan_bulk_m0_s1(qb, lpr);
            case 2: return b200_pick_scan_bulk_m1_s0(qb, lpr);
            case 3: return b200_pick_scan_bulk_m1_s1(qb, lpr);
        }
    } else {
        switch (ms) {
            case 0: return b200_pick_scan_ldg_m0_s0(qb, lpr);
            case 1: return b200_pick_scan_ldg_m0_s1(qb, lpr);
            case 2: return b200_pick_scan_ldg_m1_s0(qb, lpr);
            case 3: return b200_pick_scan_ldg_m1_s1(qb, lpr);
        }
    }
    return nullptr;
}

typedef void (*MergeFn)(const ScanParams, uint32_t);
static MergeFn pick_merge(int metric, int qb) {
#define MG(M, Q) \
    if (metric == M && qb == Q) return final_merge_kernel<M, Q>;
    MG(0, 1) MG(0, 2) MG(0, 4) MG(0, 8) MG(1, 1) MG(1, 2) MG(1, 4) MG(1, 8)
#undef MG
    return nullptr;
}

static uint32_t next_pow2(uint32_t v) {
    uint32_t m = 1;
    while (m < v) m <<= 1;
    return m;
}

static int plan_scan(b200_index* ix, int qb, int k, bool fullrank, ScanPlan* out, const ScanView* view = nullptr) {
    const ScanView v = view ? *view : index_view(ix);
    ScanPlan pl;
    pl.qb = qb;
    const int qstride = (v.d_pad + 7) / 8 * 8;
    const int kk = fullrank ? 1 : k;
    size_t budget = ix->smem_optin - 1024;
    const uint32_t step_rows = SCAN_RB * (32u / (uint32_t)v.lpr);  // rows one warp step covers
    int variant = (int)ix->opt_variant;
    // AUTO (measured over d = 64..2048, fp32 and bf16: profiles/README.md, r1_sweep11_*, r1_sweep13_*):
    // the TMA-staged ring with dynamic tiles and as many warps as fit (<= 16) x 2 stages.  Short rows
    // and bf16 rows carry more instructions per byte and want the extra warps (bf16 d=1024: 1.09-1.12
    // of the measured peak vs 1.02 for direct loads); long rows still win with 3-4 warps (d=2048: 1.12
    // vs 1.03).  Rows of 128 / 256 bytes use 8 lanes per row and the ring as well (r1_sweep14_*: d=64 fp32 1.10,
    // d=32 fp32 1.09, bf16 d=64 1.07, bf16 d=128 0.99-1.05 vs 0.82-1.00 for direct loads); other rows under 512
    // bytes go to the direct-load variant with its 32 resident warps/SM.
    const bool auto_variant = variant == B200_SCAN_AUTO;
    if (auto_variant) variant = (v.pitch >= 512 || v.lpr == 8) ? B200_VARIANT_BULK : B200_VARIANT_LDG;
    if (variant == B200_VARIANT_BULK) {
        // CTAs per SM (option scan_ctas_per_sm, default 1): c co-resident CTAs share the SM's shared memory and warp
        // slots, so that at a launch boundary (programmatic dependent launch) an SM is handed over one CTA at a time
        // instead of idling between the old CTA's exit and the new CTA's first tile.
        const int ctas = (int)std::min<int64_t>(std::max<int64_t>(ix->opt_ctas_per_sm, 1), 4);
        if (ctas > 1) budget = std::min(budget, ix->smem_per_sm / ctas - 1024 - 256);
        int nw = (int)std::min<int64_t>(std::max<int64_t>(ix->opt_warps, 1), (qb >= 4 ? 256 : B200_SCAN_THREADS_BULK) / 32 / ctas);
        nw = std::max(nw, 1);
        const int nw_min = auto_variant ? 3 : 1;
        bool ok = false;
        // Every warp owns `stages` tiles of tile_rows rows.  Prefer ~12 KB tiles, but shrink the tile
        // (down to one RB-row group) before giving up warps: 8 warps x 2 stages is what keeps the
        // ring fed.  Longer rows drop warps one at a time (7 warps at d = 896, 6 at d = 1024, 3 at
        // d = 2048 still measure 1.12-1.13); below 3 warps AUTO uses the direct-load variant instead.
        for (; nw >= nw_min && !ok; --nw) {
            uint32_t m_pref;
            if (ix->opt_tile_rows > 0)
                m_pref = (uint32_t)((ix->opt_tile_rows + step_rows - 1) / step_rows);
            else
                m_pref = (uint32_t)std::max(1.0, 12288.0 / ((double)step_rows * v.pitch) + 0.5);
            for (uint32_t m = m_pref; m >= 1 && !ok; --m) {
                uint32_t tr = m * step_rows;
                uint64_t tile_bytes = (uint64_t)tr * v.pitch;
                if (tile_bytes > (1u << 19)) continue;  // mbarrier tx-count headroom
                const uint32_t scratch = std::max<uint32_t>(B200_FINAL_BUF_KEYS, next_pow2((uint32_t)(nw * kk)));
                size_t fixed = scan_smem_bytes(B200_VARIANT_BULK, nw, qb, qstride, kk, fullrank, 0, 0, scratch);
                // ring replaces the scratch region when larger
                size_t fixed_wo_scratch = fixed - (((size_t)scratch * 8 + B200_PREF_BYTES + 127) & ~(size_t)127);
                if (fixed_wo_scratch + 64 >= budget) continue;
                size_t avail = budget - fixed_wo_scratch - 64;
                uint32_t stages = (uint32_t)std::min<uint64_t>(8, avail / ((uint64_t)nw * tile_bytes + (uint64_t)nw * 12));
                if (ix->opt_stages > 0) stages = std::min<uint32_t>(stages, (uint32_t)ix->opt_stages);
                if (stages < 2) continue;
                size_t smem = scan_smem_bytes(B200_VARIANT_BULK, nw, qb, qstride, kk, fullrank, stages, (uint32_t)tile_bytes, scratch);
                if (smem > budget) continue;
                pl.variant = B200_VARIANT_BULK;
                pl.nw = nw;
                pl.tile_rows = tr;
                pl.stages = stages;
                pl.tile_bytes = (uint32_t)tile_bytes;
                pl.scratch_keys = scratch;
                pl.smem = smem;
                pl.grid = ix->num_sms * ctas;  // persistent CTAs
                ok = true;
            }
        }
        if (!ok) variant = B200_VARIANT_LDG;
    }
    if (variant == B200_VARIANT_LDG) {
        int nw = (int)std::min<int64_t>(std::max<int64_t>(ix->opt_warps, 1), B200_SCAN_THREADS_LDG / 32);
        for (;; nw >>= 1) {
            if (nw < 1) return fail("k=%d does not fit the fused top-k shared-memory budget", k);
            const uint32_t scratch = std::max<uint32_t>(B200_FINAL_BUF_KEYS, next_pow2((uint32_t)(nw * kk)));
            size_t smem = scan_smem_bytes(B200_VARIANT_LDG, nw, qb, qstride, kk, fullrank, 0, 0, scratch);
            if (smem > budget) continue;
            pl.variant = B200_VARIANT_LDG;
            pl.nw = nw;
            pl.tile_rows = step_rows;
            pl.stages = 0;
            pl.tile_bytes = 0;
            pl.scratch_keys = scratch;
            pl.smem = smem;
            break;
        }
        ScanFn fn = pick_scan(v.metric, v.store, qb, B200_VARIANT_LDG, v.lpr);
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, pl.nw * 32, pl.smem));
        if (occ < 1) return fail("LDG scan kernel does not fit on an SM (smem %zu)", pl.smem);
        if (ix->opt_ctas_per_sm > 0) occ = (int)std::min<int64_t>(occ, ix->opt_ctas_per_sm);
        pl.grid = occ * ix->num_sms;
    }
    *out = pl;
    return 0;
}

template <int METRIC>
__global__ void fill_pad_kernel(float* D, int64_t* I, int64_t count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) {
        D[i] = METRIC == 0 ? -FLT_MAX : FLT_MAX;
        I[i] = -1;
    }
}

static int launch_scan(b200_index* ix, const ScanPlan& pl, const float* q_dev, int nqb, int k, float* D,
                       int64_t* I, uint32_t* score_keys, cudaStream_t st, const ScanView* view = nullptr) {
    const ScanView v = view ? *view : index_view(ix);
    ScanParams p;
    memset(&p, 0, sizeof p);
    p.rows = v.rows;
    p.pitch_bytes = v.pitch;
    p.nvec = (uint32_t)(v.pitch / 16);
    p.n = (uint64_t)ix->ntotal;
    p.q = q_dev;
    p.d = v.d;
    p.qstride = (v.d_pad + 7) / 8 * 8;
    p.nqb = nqb;
    p.k = score_keys ? 1 : k;
    p.normalize_q = v.normalize_q ? 1 : 0;
    const int set = (int)(ix->launch_seq++ & 1);  // control words and survivor lists of this launch's parity
    p.ticket = reinterpret_cast<unsigned int*>(ix->ctl + (size_t)set * 16);
    p.dynamic = ix->opt_dynamic < 0 ? (pl.variant == B200_VARIANT_BULK ? 1 : 0) : (ix->opt_dynamic ? 1 : 0);
    {
        // One atomic claim hands out a run of tiles.  The run bounds the tail (a warp finishes at
        // most one run after the database is exhausted: run x ~2 us) while a single hot counter
        // sustains only ~400 claims/us; >= 64 claims per warp, runs of 4..16 tiles that shrink to
        // claim_min towards the end (guided self-scheduling in the kernel).
        uint64_t tiles = ((uint64_t)ix->ntotal + pl.tile_rows - 1) / pl.tile_rows;
        uint64_t per = tiles / ((uint64_t)pl.grid * pl.nw * 64);
        p.claim_chunk = (uint32_t)std::min<uint64_t>(std::max<uint64_t>(per, 4), 16);
        if (ix->opt_claim_chunk > 0) p.claim_chunk = (uint32_t)ix->opt_claim_chunk;
        p.claim_min = ix->opt_claim_min > 0 ? (uint32_t)ix->opt_claim_min : 2u;
        if (p.claim_min > p.claim_chunk) p.claim_min = p.claim_chunk;
        // the first run of every warp is static; small databases spread their tiles over all warps
        const uint64_t warps = (uint64_t)pl.grid * pl.nw;
        p.claim_first = (uint32_t)std::min<uint64_t>(p.claim_chunk, std::max<uint64_t>(1, tiles / warps));
        if (ix->opt_claim_first > 0) p.claim_first = (uint32_t)std::min<uint64_t>((uint64_t)ix->opt_claim_first, std::max<uint64_t>(1, tiles / warps));
    }
    p.fused_tail = (nqb == 1 || score_keys) ? 1 : 0;
    if (ix->opt_fused_tail >= 0) p.fused_tail = ix->opt_fused_tail ? 1 : 0;
    p.D = D;
    p.I = I;
    p.id_map = v.id_map;
    p.id_base = 0;
    p.tile_rows = pl.tile_rows;
    p.stages = pl.stages;
    p.tile_bytes = pl.tile_bytes;
    p.evict_first = (int)ix->opt_evict_first;
    p.score_keys = score_keys;
    p.row_mask = ix->cur_mask;
    p.scratch_keys = pl.scratch_keys;
    if (ix->xchg_active && !score_keys) {
        p.xchg_peers = ix->xchg_peers_dev;
        p.xchg_world = ix->xchg_world;
        p.xchg_rank = ix->xchg_rank;
        p.xchg_epoch = ++ix->xchg_epoch;
        p.xchg_slot_bytes = (uint32_t)XCHG_SLOT_BYTES;
        p.xchg_status = ix->xchg_status;
        p.fused_tail = 1;  // the exchange lives in the last CTA's tail
    }
    if (pl.grid > 1024) return fail("scan grid of %d CTAs exceeds the final merge's selector table", pl.grid);
    if (!score_keys) {
        const size_t need = (size_t)pl.grid * pl.qb * k;
        if (ix->partials_cap < need) {
            CK(cudaStreamSynchronize(st));  // launches in flight still use the old buffer
            if (ix->partials) CK(cudaFree(ix->partials));
            ix->partials = nullptr;
            ix->partials_cap = 0;
            CK(cudaMalloc((void**)&ix->partials, 2 * need * sizeof(uint64_t)));
            ix->partials_cap = need;
        }
        p.partials = ix->partials + (size_t)set * ix->partials_cap;
    }
    if (ix->opt_phase_stamps) {
        const size_t need = (size_t)pl.grid * 8;
        if (ix->stamps_cap < need) {
            CK(cudaStreamSynchronize(st));
            CKI(grow(&ix->stamps, &ix->stamps_cap, need));
        }
        CK(cudaMemsetAsync(ix->stamps, 0, need * sizeof(unsigned long long), st));
        p.stamps = ix->stamps;
        ix->stamps_grid = pl.grid;
    }
    ScanFn fn = pick_scan(v.metric, v.store, pl.qb, pl.variant, v.lpr);
    if (!fn) return fail("no scan kernel for metric=%d store=%d qb=%d variant=%d", v.metric, v.store, pl.qb, pl.variant);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    // Programmatic dependent launch: a search that follows another kernel on this stream may start while that kernel
    // drains (its CTAs take SMs as they free up).  pdl = 1: the kernel waits for its predecessor before it reads the
    // queries; pdl = 2 (option queries_stable: the caller promises that the queries are not produced by the preceding
    // kernel on the stream): it waits only after its scan, before it publishes anything.
    p.pdl = ix->opt_pdl ? (ix->opt_queries_stable ? 2 : 1) : 0;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.gridDim = dim3((unsigned)pl.grid);
    cfg.blockDim = dim3((unsigned)(pl.nw * 32));
    cfg.dynamicSmemBytes = pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = p.pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, fn, p));
    ++ix->launches;
    if (!score_keys && !p.fused_tail) {
        MergeFn mf = pick_merge(v.metric, pl.qb);
        size_t msmem = (size_t)pl.scratch_keys * 8 + 16 + B200_PREF_BYTES;
        mf<<<nqb, 256, msmem, st>>>(p, (uint32_t)pl.grid);
        ++ix->launches;
        CK(cudaGetLastError());
    }
    return 0;
}

// Full ranking of a SMALL database (memo-sized: n <= 4096 rows; at 10k rows the single-CTA sort already loses to the radix passes) in one launch: one CTA per query builds the 64-bit
// (score key, row) keys in shared memory, sorts them with the CTA bitonic network and emits the first k —
// instead of the 14 launches of the radix path (memo's k = ntotal call at 1k rows: ~100 us -> one scan + one sort).
// The 64-bit keys carry the tie rule, so the order equals the stable radix sort's.
#define FULLRANK_SMALL_MAX 4096
template <int METRIC>
__global__ void __launch_bounds__(1024) fullrank_small_kernel(const uint32_t* __restrict__ hi_keys, uint32_t n, uint32_t m,
                                                              int64_t k, const int64_t* __restrict__ id_map,
                                                              float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ uint64_t fr_keys[];
    const uint32_t* hi = hi_keys + (size_t)blockIdx.x * n;
    for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
        uint32_t h = i < n ? hi[i] : 0u;
        fr_keys[i] = h ? (((uint64_t)h << 32) | (uint64_t)(0xFFFFFFFFu - i)) : 0ull;
    }
    cta_bitonic_sort_desc(fr_keys, m);
    float* Dq = D + (size_t)blockIdx.x * k;
    int64_t* Iq = I + (size_t)blockIdx.x * k;
    for (int64_t i = threadIdx.x; i < k; i += blockDim.x) {
        const uint64_t key = i < (int64_t)m ? fr_keys[i] : 0ull;
        float dist = (METRIC == 0) ? -FLT_MAX : FLT_MAX;
        int64_t id = -1;
        if (key != 0ull) {
            dist = b200_key_score(key, METRIC);
            const uint32_t row = b200_key_row(key);
            id = id_map ? id_map[row] : (int64_t)row;
        }
        Dq[i] = dist;
        Iq[i] = id;
    }
}

static int fullrank_one(b200_index* ix, const uint32_t* hi_keys, int64_t k, float* D, int64_t* I, cudaStream_t st) {
    const uint64_t n = (uint64_t)ix->ntotal;
    const uint32_t nblocks = (uint32_t)((n + RADIX_CHUNK - 1) / RADIX_CHUNK);
    if (ix->fr_cap < n) {
        for (int i = 0; i < 4; ++i) {
            if (ix->fr_buf[i]) CK(cudaFree(ix->fr_buf[i]));
            ix->fr_buf[i] = nullptr;
        }
        ix->fr_cap = 0;
        for (int i = 0; i < 4; ++i) CK(cudaMalloc((void**)&ix->fr_buf[i], n * sizeof(uint32_t)));
        ix->fr_cap = n;
    }
    CKI(grow(&ix->fr_hist, &ix->fr_hist_cap, (size_t)256 * nblocks));
    uint32_t *ka = ix->fr_buf[0], *kb = ix->fr_buf[1], *va = ix->fr_buf[2], *vb = ix->fr_buf[3];
    fullrank_prepare_kernel<<<ix->num_sms * 4, 256, 0, st>>>(hi_keys, ka, va, n);
    ++ix->launches;
    for (int pass = 0; pass < 4; ++pass) {
        int shift = 8 * pass;
        radix_hist_kernel<<<nblocks, RADIX_THREADS, 0, st>>>(ka, n, shift, ix->fr_hist, nblocks);
        radix_scan_kernel<<<1, 1024, 0, st>>>(ix->fr_hist, (uint64_t)256 * nblocks);
        radix_scatter_kernel<<<nblocks, RADIX_THREADS, 0, st>>>(ka, va, kb, vb, n, shift, ix->fr_hist, nblocks);
        ix->launches += 3;
        std::swap(ka, kb);
        std::swap(va, vb);
    }
    const int64_t* idm = ix->ids_state == 1 ? ix->ids : nullptr;
    unsigned eb = (unsigned)std::min<int64_t>((k + 255) / 256, (int64_t)ix->num_sms * 8);
    if (ix->metric == B200_METRIC_IP)
        fullrank_emit_kernel<0><<<eb, 256, 0, st>>>(ka, va, n, k, idm, 0, D, I);
    else
        fullrank_emit_kernel<1><<<eb, 256, 0, st>>>(ka, va, n, k, idm, 0, D, I);
    ++ix->launches;
    CK(cudaGetLastError());
    return 0;
}

static int pick_qb(int64_t remaining, int64_t forced, int lpr) {
    if (lpr != 32) return remaining >= 2 ? 8 : 1;  // short-row kernels exist for query blocks 1 and 8
    if (forced == 1 || forced == 2 || forced == 4 || forced == 8) return (int)forced;
    if (remaining >= 8) return 8;
    if (remaining > 2) return 4;
    if (remaining == 2) return 2;
    return 1;
}


// ---------------------------------------------------------------------------------------------
// K3: batched search on the tensor cores (gemm_topk.cuh)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int get_encode_fn(EncodeTiledFn* out) {
    static EncodeTiledFn cached = nullptr;
    if (!cached) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (qres != cudaDriverEntryPointSuccess || !fn) return fail("cuTensorMapEncodeTiled is not available from the driver");
        cached = (EncodeTiledFn)fn;
    }
    *out = cached;
    return 0;
}
// bf16 K-major matrix [rows, kpad] -> 2-D tensor map with a {64, box_rows} box and 128-byte swizzle
static int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint32_t kpad, uint32_t box_rows) {
    EncodeTiledFn enc = nullptr;
    CKI(get_encode_fn(&enc));
    cuuint64_t gdim[2] = {kpad, rows};
    cuuint64_t gstride[1] = {(cuuint64_t)kpad * 2};
    cuuint32_t box[2] = {G3_BLOCK_K, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

static bool gemm_eligible(b200_index* ix, int64_t nq, int64_t k) {
    return ix->opt_gemm_min_nq > 0 && nq >= ix->opt_gemm_min_nq && k <= B200_FUSED_K_MAX &&
           ix->ntotal >= ix->opt_gemm_min_rows && ix->ntotal >= 512 * k && ix->d >= 32 && k < ix->opt_fullrank_min_k;
}

// K of the shadow GEMM: d (+2 columns carrying |y|^2/2 and -1 for L2), padded to whole 64-column blocks
static int gemm_kpad(const b200_index* ix) {
    const int cols = ix->d + (ix->metric == B200_METRIC_L2 ? 2 : 0);
    return (cols + G3_BLOCK_K - 1) / G3_BLOCK_K * G3_BLOCK_K;
}

// Sampled tiles the threshold pass needs at least: theta is never taken above the 8th largest sampled group maximum,
// so a query emits about 8 * n / sampled_rows candidates however few are wanted; keeping that under a quarter of the
// candidate list (1024 of 4096) takes a sample that grows with the database (100M rows: 3052 tiles = 0.8 % of it).
static uint32_t gemm_min_sample_tiles(int64_t n) {
    const double t = std::ceil(8.0 * (double)n / (1024.0 * G3_BLOCK_N));
    return (uint32_t)std::min(8192.0, std::max(1.0, t));
}
// queries per K3 call: bounds the candidate lists (16 KB per query) and the sampled maxima (32 B per query and tile)
static int64_t gemm_query_block(const b200_index* ix) {
    const uint32_t t = std::max<uint32_t>(1024, gemm_min_sample_tiles(ix->ntotal));
    return std::max<int64_t>(1024, (int64_t)16384 * 1024 / t / 256 * 256);
}

// Rows of the streamed scratch (a multiple of 512: two chunk buffers of whole 256-row tiles): two chunks of ~256 MB
// of bf16 each — large enough to amortise the two launches per chunk, and the converter of chunk c+1 overlaps the sweep
// of chunk c — but at least the sampled tiles of the threshold pass, which are packed into the same scratch.
static size_t shadow_stream_rows(const b200_index* ix, int kpad) {
    size_t rows = 2 * (((size_t)256 << 20) / ((size_t)kpad * 2) / G3_BLOCK_N * G3_BLOCK_N);
    rows = std::max<size_t>(rows, ((size_t)std::max<uint32_t>(1024, gemm_min_sample_tiles(ix->ntotal)) + 1) / 2 * 2 * G3_BLOCK_N);
    if (ix->opt_gemm_shadow_max_rows > 0)
        rows = std::max<size_t>(2 * G3_BLOCK_N, (size_t)ix->opt_gemm_shadow_max_rows / (2 * G3_BLOCK_N) * (2 * G3_BLOCK_N));
    return rows;
}

// 0: the whole bf16 shadow is resident and current.  3: it does not fit next to the rows (or option
// gemm_shadow_max_rows caps it): sh_rows is a scratch of sh_cap_rows rows and the search streams the rows through it
// chunk by chunk.  2: not even the scratch could be allocated — the caller uses the scan path.
static int ensure_shadow(b200_index* ix, cudaStream_t st) {
    const int kpad = gemm_kpad(ix);
    if (!ix->sh_maxnorm) CK(cudaMalloc((void**)&ix->sh_maxnorm, sizeof(unsigned int)));
    const bool capped = ix->opt_gemm_shadow_max_rows > 0 && ix->ntotal > ix->opt_gemm_shadow_max_rows;
    if (!ix->sh_streamed && !capped && ix->sh_valid_rows == ix->ntotal) return 0;
    if (ix->sh_streamed && ix->sh_valid_rows == ix->ntotal && ix->sh_cap_rows == shadow_stream_rows(ix, kpad)) return 3;
    bool want_stream = capped;
    if (!want_stream && (ix->sh_streamed || ix->sh_cap_rows < (size_t)ix->ntotal)) {
        CK(cudaStreamSynchronize(st));
        if (ix->sh_rows) CK(cudaFree(ix->sh_rows));
        ix->sh_rows = nullptr;
        ix->sh_cap_rows = 0;
        ix->sh_streamed = false;
        size_t cap = (size_t)ix->ntotal;  // the rows present, not the reserved capacity: the shadow is rebuilt on change anyway
        cudaError_t e = cudaMalloc((void**)&ix->sh_rows, cap * kpad * sizeof(__nv_bfloat16));
        if (e != cudaSuccess) {
            cudaGetLastError();
            ix->sh_rows = nullptr;
            want_stream = true;  // no room for a resident copy: stream the rows through a scratch instead
        } else {
            ix->sh_cap_rows = cap;
        }
    }
    if (want_stream) {
        const size_t rows = shadow_stream_rows(ix, kpad);
        if (!ix->sh_streamed || ix->sh_cap_rows != rows) {
            CK(cudaStreamSynchronize(st));
            if (ix->sh_rows) CK(cudaFree(ix->sh_rows));
            ix->sh_rows = nullptr;
            ix->sh_cap_rows = 0;
            ix->sh_streamed = false;
            cudaError_t e = cudaMalloc((void**)&ix->sh_rows, rows * kpad * sizeof(__nv_bfloat16));
            if (e != cudaSuccess) {
                cudaGetLastError();
                ix->sh_rows = nullptr;
                ix->sh_failed_rows = ix->ntotal;
                fail("no room for the %.2f GB bf16 scratch: %s", (double)rows * kpad * 2 / 1e9, cudaGetErrorString(e));
                return 2;
            }
            ix->sh_cap_rows = rows;
            ix->sh_streamed = true;
        }
        ix->sh_valid_rows = ix->ntotal;
        return 3;
    }
    CK(cudaMemsetAsync(ix->sh_maxnorm, 0, sizeof(unsigned int), st));
    shadow_rows_kernel<<<ix->num_sms * 4, 256, 0, st>>>(ix->rows, ix->pitch, ix->store, (uint64_t)ix->ntotal, (uint64_t)ix->ntotal, 1,
                                                        ix->d, kpad, ix->metric == B200_METRIC_L2 ? 1 : 0, ix->sh_rows, nullptr,
                                                        ix->sh_maxnorm);
    ++ix->launches;
    CK(cudaGetLastError());
    ix->sh_valid_rows = ix->ntotal;
    return 0;
}

static int search_scan_block(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev,
                             cudaStream_t st);

// bound_out != nullptr selects the row-sharded form (b200_index_search_shard_dev): this index is one of `world`
// shards, thresholds aim at 1/world of the candidates, the re-rank leaves each query's exclusion bound in
// bound_out[nq] and nothing is read back — the certificate is taken after the merge over all shards.
static int search_gemm(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev,
                       cudaStream_t st, int depth = 0, float* bound_out = nullptr, int world = 1) {
    const int kpad = gemm_kpad(ix);
    const uint64_t n = (uint64_t)ix->ntotal;
    const uint32_t NT = (uint32_t)((n + G3_BLOCK_N - 1) / G3_BLOCK_N);
    const uint32_t m_tiles = (uint32_t)((nq + G3_BLOCK_M - 1) / G3_BLOCK_M);
    const uint32_t cap = 4096;
    const int cg = ix->opt_gemm_cta_group == 1 ? 1 : 2;
    const auto host_t0 = std::chrono::steady_clock::now();
    const int shadow_state = ensure_shadow(ix, st);  // 2 = not even a scratch: the caller uses the scan path
    if (shadow_state != 0 && shadow_state != 3) return shadow_state;
    const bool streamed = shadow_state == 3;
    const size_t S = ix->sh_cap_rows;  // streamed: rows per chunk
    cudaEvent_t* ev = ix->g_ev;
    for (int i = 0; i < 5; ++i)
        if (!ev[i]) CK(cudaEventCreate(&ev[i]));
    CK(cudaEventRecord(ev[4], st));
    // Small batches: rows as the M operand, the queries resident in shared memory (gemm_rows_topk_kernel) when
    // N = nq rounded up to 16 fits next to a ring of >= 4 stages.
    const uint32_t n_cols = (uint32_t)((nq + 15) / 16 * 16);
    uint32_t rows_stages = 0;
    size_t rows_smem = 0;
    if (cg == 2 && ix->opt_gemm_rows_form != 0 && n_cols <= 256) {
        const size_t qres = (size_t)(n_cols / 2) * kpad * 2;
        const size_t budget = ix->smem_optin - 1024 - 256 - (size_t)((n_cols + 31) & ~31u) * 4;
        if (qres + 4 * (size_t)G3T_A_BYTES <= budget) {
            rows_stages = (uint32_t)std::min<size_t>(8, (budget - qres) / G3T_A_BYTES);
            rows_smem = 1024 + qres + (size_t)rows_stages * G3T_A_BYTES + (2 * rows_stages + 6) * 8 + (size_t)((n_cols + 31) & ~31u) * 4;
        }
    }
    const bool rows_form = rows_stages >= 4;
    ix->stat_gemm_rows_form = rows_form ? 1 : 0;
    // ---- scratch ----
    const size_t qb_elems = (size_t)((m_tiles + 1) / 2 * 2) * G3_BLOCK_M * kpad;  // whole 256-query groups
    // sampled tiles: each contributes 8 group maxima; at most 8192 maxima per query are sorted
    // expected emissions per query = emit_factor * k over ALL shards; a retry of uncertified queries (depth 1)
    // widens the net 3x
    double want = (double)std::max<int64_t>(ix->opt_gemm_emit_factor, 2) * (depth ? 3.0 : 1.0) * (double)k;
    want = std::min(want, 0.7 * cap);
    // (a shard aims 30% above its share: the merged certificate has to beat the HIGHEST of the shards' thresholds,
    // which independent estimates push up)
    if (world > 1) want = std::max(1.3 * want / world, 16.0);
    uint32_t T = (uint32_t)std::min<int64_t>(std::min<int64_t>(std::max<int64_t>(ix->opt_gemm_sample_tiles, 64), 1024), NT);
    T = std::min<uint32_t>(NT, std::max<uint32_t>(T, gemm_min_sample_tiles(ix->ntotal)));
    if (streamed) T = (uint32_t)std::min<size_t>(T, S / G3_BLOCK_N);  // the sampled tiles are packed into the scratch
    if (world > 1) {
        // a shard samples only as many tiles as keep the selected rank near 10 (the rank is want * sampled / n)
        const double t_need = 10.5 * (double)n / (want * G3_BLOCK_N);
        T = (uint32_t)std::min<double>(T, std::max(64.0, std::ceil(t_need)));
    }
    const uint32_t stride = std::max<uint32_t>(1, NT / T);
    if (ix->g_qb_cap < qb_elems || ix->g_q_cap < (size_t)nq || ix->g_tilemax_cap < (size_t)nq * T * 8 ||
        ix->g_cand_cap < (size_t)nq * cap)
        CK(cudaStreamSynchronize(st));
    CKI(grow(&ix->g_qb, &ix->g_qb_cap, qb_elems));
    if (ix->g_q_cap < (size_t)nq) {
        cudaFree(ix->g_qnorm2); cudaFree(ix->g_theta); cudaFree(ix->g_count); cudaFree(ix->g_cert);
        ix->g_qnorm2 = nullptr; ix->g_theta = nullptr; ix->g_count = nullptr; ix->g_cert = nullptr;
        ix->g_q_cap = 0;
        CK(cudaMalloc((void**)&ix->g_qnorm2, (size_t)nq * 4));
        CK(cudaMalloc((void**)&ix->g_theta, (size_t)nq * 4));
        CK(cudaMalloc((void**)&ix->g_count, (size_t)nq * 4));
        CK(cudaMalloc((void**)&ix->g_cert, (size_t)nq * 4));
        ix->g_q_cap = (size_t)nq;
    }
    CKI(grow(&ix->g_tilemax, &ix->g_tilemax_cap, (size_t)nq * T * 8));
    CKI(grow(&ix->g_cand, &ix->g_cand_cap, (size_t)nq * cap));
    // ---- query shadow (zero padded to whole 128-query tiles) ----
    CK(cudaMemsetAsync(ix->g_qb, 0, qb_elems * sizeof(__nv_bfloat16), st));
    shadow_rows_kernel<<<(unsigned)std::min<int64_t>((nq + 7) / 8, ix->num_sms * 8), 256, 0, st>>>(
        (const uint8_t*)q_dev, (uint64_t)ix->d * 4, 0, (uint64_t)nq, (uint64_t)nq, 1, ix->d, kpad, ix->metric == B200_METRIC_L2 ? 2 : 0,
        ix->g_qb, ix->g_qnorm2, nullptr);
    ++ix->launches;
    CK(cudaGetLastError());
    CUtensorMap tm_q, tm_db;
    CKI(make_tmap_bf16(&tm_q, ix->g_qb, (uint64_t)((m_tiles + 1) / 2 * 2) * G3_BLOCK_M, (uint32_t)kpad, G3_BLOCK_M));
    CKI(make_tmap_bf16(&tm_db, ix->sh_rows, streamed ? (uint64_t)S : n, (uint32_t)kpad, G3_BLOCK_N / cg));
    const size_t C2 = S / 2;  // streamed: rows per chunk buffer
    CUtensorMap tm_half[2];
    if (streamed)
        for (int b = 0; b < 2; ++b)
            CKI(make_tmap_bf16(&tm_half[b], ix->sh_rows + (size_t)b * C2 * kpad, (uint64_t)C2, (uint32_t)kpad, G3_BLOCK_N / cg));
    const bool masked = ix->cur_mask != nullptr;
    CUtensorMap tm_q_rows;
    if (rows_form) CKI(make_tmap_bf16(&tm_q_rows, ix->g_qb, (uint64_t)((m_tiles + 1) / 2 * 2) * G3_BLOCK_M, (uint32_t)kpad, n_cols / 2));
    typedef void (*GemmRowsFn)(const CUtensorMap, const CUtensorMap, const GemmRowsParams);
    GemmRowsFn rfns[2] = {masked ? gemm_rows_topk_kernel<G3_MODE_TILEMAX, true> : gemm_rows_topk_kernel<G3_MODE_TILEMAX, false>,
                          masked ? gemm_rows_topk_kernel<G3_MODE_EMIT, true> : gemm_rows_topk_kernel<G3_MODE_EMIT, false>};
    if (rows_form)
        for (GemmRowsFn f : rfns) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rows_smem));
    typedef void (*GemmFn)(const CUtensorMap, const CUtensorMap, const GemmParams);
    // [mode]: the tile-maxima pass and the emit pass are separate instantiations (no mode branches in the epilogue)
    GemmFn gfns[2] = {
        cg == 1 ? (masked ? gemm_topk_kernel<1, true, G3_MODE_TILEMAX> : gemm_topk_kernel<1, false, G3_MODE_TILEMAX>)
                : (masked ? gemm_topk_kernel<2, true, G3_MODE_TILEMAX> : gemm_topk_kernel<2, false, G3_MODE_TILEMAX>),
        cg == 1 ? (masked ? gemm_topk_kernel<1, true, G3_MODE_EMIT> : gemm_topk_kernel<1, false, G3_MODE_EMIT>)
                : (masked ? gemm_topk_kernel<2, true, G3_MODE_EMIT> : gemm_topk_kernel<2, false, G3_MODE_EMIT>)};
    const size_t gsmem = cg == 1 ? G3Cfg<1>::kSmemBytes : G3Cfg<2>::kSmemBytes;
    for (GemmFn f : gfns) CK(cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem));
    auto launch_gemm = [&](const GemmParams& g, const CUtensorMap& tm_rows) -> int {
        if (rows_form) {
            GemmRowsParams r;
            memset(&r, 0, sizeof r);
            r.nq = g.nq; r.n_cols = n_cols; r.n = g.n; r.k_blocks = g.k_blocks;
            r.tile_first = g.tile_first; r.tile_stride = g.tile_stride; r.tile_count = g.tile_count;
            r.src_tile_first = g.src_tile_first; r.src_tile_stride = g.src_tile_stride;
            r.stages = rows_stages; r.mode = g.mode; r.theta = g.theta; r.cand_count = g.cand_count; r.cand_rows = g.cand_rows;
            r.cand_cap = g.cand_cap; r.tilemax = g.tilemax; r.row_mask = g.row_mask;
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)(ix->num_sms / 2 * 2));
            cfg.blockDim = dim3(G3T_THREADS);
            cfg.dynamicSmemBytes = rows_smem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, rfns[g.mode == G3_MODE_EMIT ? 1 : 0], tm_q_rows, tm_rows, r));
            ++ix->launches;
            CK(cudaGetLastError());
            return 0;
        }
        GemmFn gfn = gfns[g.mode == G3_MODE_EMIT ? 1 : 0];
        if (cg == 1) {
            gfn<<<ix->num_sms, G3_THREADS, gsmem, st>>>(tm_q, tm_rows, g);
        } else {
            cudaLaunchConfig_t cfg;
            memset(&cfg, 0, sizeof cfg);
            cfg.gridDim = dim3((unsigned)(ix->num_sms / 2 * 2));
            cfg.blockDim = dim3(G3_THREADS);
            cfg.dynamicSmemBytes = gsmem;
            cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, gfn, tm_q, tm_rows, g));
        }
        ++ix->launches;
        CK(cudaGetLastError());
        return 0;
    };
    GemmParams gp;
    memset(&gp, 0, sizeof gp);
    gp.nq = (uint32_t)nq;
    gp.m_tiles = m_tiles;
    gp.n = n;
    gp.k_blocks = (uint32_t)(kpad / G3_BLOCK_K);
    {
        // Work units that run together share a chunk of row tiles through L2: keep the concurrently
        // streamed chunks within ~48 MB of the 126 MB L2.
        const uint32_t groups = (uint32_t)(ix->num_sms / cg);
        const uint32_t m_groups = (m_tiles + cg - 1) / cg;
        const uint32_t concurrent = (groups + m_groups - 1) / m_groups + 1;
        const double tile_bytes = (double)G3_BLOCK_N * kpad * 2;
        double ct = 48e6 / (concurrent * tile_bytes);
        gp.chunk_tiles = (uint32_t)std::min(64.0, std::max(4.0, ct));
        if (ix->opt_gemm_chunk_tiles > 0) gp.chunk_tiles = (uint32_t)ix->opt_gemm_chunk_tiles;
    }
    gp.theta = ix->g_theta;
    gp.cand_count = ix->g_count;
    gp.cand_rows = ix->g_cand;
    gp.cand_cap = cap;
    gp.tilemax = ix->g_tilemax;
    gp.row_mask = ix->cur_mask;  // filtered search: excluded rows never become candidates
    // ---- pass 1: tile maxima over a strided sample of T tiles -> theta ----
    gp.mode = G3_MODE_TILEMAX;
    gp.tile_first = 0;
    gp.tile_stride = stride;
    gp.tile_count = T;
    gp.src_tile_first = 0;
    gp.src_tile_stride = stride;
    const int aug_rows = ix->metric == B200_METRIC_L2 ? 1 : 0;
    CK(cudaEventRecord(ev[0], st));
    if (streamed) {
        // the sampled tiles, packed back to back into the scratch; the maximum row norm is re-collected by the
        // chunks of pass 2, which convert every row
        CK(cudaMemsetAsync(ix->sh_maxnorm, 0, sizeof(unsigned int), st));
        shadow_rows_kernel<<<ix->num_sms * 4, 256, 0, st>>>(ix->rows, ix->pitch, ix->store, (uint64_t)T * G3_BLOCK_N, n, stride, ix->d, kpad,
                                                            aug_rows, ix->sh_rows, nullptr, ix->sh_maxnorm);
        ++ix->launches;
        CK(cudaGetLastError());
        gp.src_tile_stride = 1;
    }
    CKI(launch_gemm(gp, tm_db));
    {
        // expected emissions per query = emit_factor * k.  The rank-th largest of the sampled
        // 32-row group maxima estimates the score quantile (sample rows / n) * that count; the rank
        // is kept below G/8 so that two of the top scores rarely share a group.
        const uint32_t G = T * 8;
        double sample_rows = (double)T * G3_BLOCK_N;
        double r = want * std::min(1.0, sample_rows / (double)n);
        uint32_t rank = (uint32_t)std::min<double>(std::max(r, 8.0), (double)(G / 8));
        select_theta_kernel<<<(unsigned)nq, 256, 0, st>>>(ix->g_tilemax, G, rank, ix->g_theta);
        ++ix->launches;
        CK(cudaGetLastError());
    }
    // ---- pass 2: emit candidates over every tile ----
    CK(cudaMemsetAsync(ix->g_count, 0, (size_t)nq * 4, st));
    gp.mode = G3_MODE_EMIT;
    gp.tile_stride = 1;
    gp.src_tile_stride = 1;
    CK(cudaEventRecord(ev[1], st));
    if (!streamed) {
        gp.tile_first = gp.src_tile_first = 0;
        gp.tile_count = NT;
        CKI(launch_gemm(gp, tm_db));
    } else {
        // streamed shadow: the fp32 rows are rounded chunk by chunk into the two halves of the scratch; the converter
        // runs on a second stream, so chunk c+1 is converted (HBM-bound) while the GEMM sweeps chunk c
        if (!ix->sh_stream2) CK(cudaStreamCreateWithFlags(&ix->sh_stream2, cudaStreamNonBlocking));
        for (int i = 0; i < 5; ++i)
            if (!ix->sh_ev[i]) CK(cudaEventCreateWithFlags(&ix->sh_ev[i], cudaEventDisableTiming));
        cudaStream_t s2 = ix->sh_stream2;
        CK(cudaEventRecord(ix->sh_ev[4], st));  // fork: the threshold pass (which used the whole scratch) is enqueued
        CK(cudaStreamWaitEvent(s2, ix->sh_ev[4], 0));
        uint64_t c = 0;
        for (uint64_t r0 = 0; r0 < n; r0 += C2, ++c) {
            const int b = (int)(c & 1);
            const uint64_t rows_c = std::min<uint64_t>(C2, n - r0);
            if (c >= 2) CK(cudaStreamWaitEvent(s2, ix->sh_ev[2 + b], 0));  // the sweep that last read this half is done
            // 2 CTAs of 256 threads x 64 registers per SM (16 warps with 4 KB in flight each): the GEMM's CTA (192 threads
            // x 120 registers, one per SM) must find thread slots and registers beside them
            shadow_rows_kernel<<<ix->num_sms * 2, 256, 0, s2>>>(ix->rows + r0 * ix->pitch, ix->pitch, ix->store, rows_c, rows_c, 1, ix->d,
                                                                kpad, aug_rows, ix->sh_rows + (size_t)b * C2 * kpad, nullptr, ix->sh_maxnorm);
            ++ix->launches;
            CK(cudaGetLastError());
            CK(cudaEventRecord(ix->sh_ev[b], s2));
            CK(cudaStreamWaitEvent(st, ix->sh_ev[b], 0));
            gp.tile_first = (uint32_t)(r0 / G3_BLOCK_N);
            gp.src_tile_first = 0;
            gp.tile_count = (uint32_t)((rows_c + G3_BLOCK_N - 1) / G3_BLOCK_N);
            CKI(launch_gemm(gp, tm_half[b]));
            CK(cudaEventRecord(ix->sh_ev[2 + b], st));
        }
    }
    ix->stat_gemm_streamed = streamed ? 1 : 0;
    CK(cudaEventRecord(ev[2], st));
    // ---- exact re-rank + certificate ----
    RerankParams rp;
    memset(&rp, 0, sizeof rp);
    rp.rows = ix->rows;
    rp.pitch_bytes = ix->pitch;
    rp.nvec = (uint32_t)(ix->pitch / 16);
    rp.store = ix->store;
    rp.d = ix->d;
    rp.qstride = (ix->d_pad + 7) / 8 * 8;
    rp.lpr = ix->lpr;
    rp.q = q_dev;
    rp.cand_rows = ix->g_cand;
    rp.cand_count = ix->g_count;
    rp.cand_cap = cap;
    rp.theta = ix->g_theta;
    rp.qnorm2 = ix->g_qnorm2;
    rp.max_norm2_bits = ix->sh_maxnorm;
    // |approx - exact| <= (2u + u^2) |q||y| with u = 2^-8 (two bf16 roundings), plus fp32 accumulation
    // slack on both sides (d * 2^-22 each, generous)
    rp.eps_rel = 0.0078278f + 2.0f * (float)ix->d * 2.4e-7f + 1e-4f;
    rp.n = n;
    rp.k = (int)k;
    rp.id_map = ix->ids_state == 1 ? ix->ids : nullptr;
    rp.D = D_dev;
    rp.I = I_dev;
    rp.certified = ix->g_cert;
    rp.bound = bound_out;
    const size_t rsmem = (size_t)rp.qstride * 4 + (size_t)cap * 8;
    CK(cudaFuncSetAttribute(rerank_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    CK(cudaFuncSetAttribute(rerank_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
    if (ix->metric == B200_METRIC_IP)
        rerank_kernel<0><<<(unsigned)nq, 256, rsmem, st>>>(rp);
    else
        rerank_kernel<1><<<(unsigned)nq, 256, rsmem, st>>>(rp);
    ++ix->launches;
    CK(cudaGetLastError());
    CK(cudaEventRecord(ev[3], st));
    if (bound_out) {  // row-sharded form: nothing is read back here
        ix->stat_gemm_used = 1;
        ix->stat_gemm_fallbacks = -1;
        ix->stat_gemm_cand_total = -1;
        ix->g_ev_pending = true;
        return 0;
    }
    // ---- uncertified queries go through the exact scan ----
    std::vector<int> cert((size_t)nq);
    std::vector<unsigned int> counts((size_t)nq);
    CK(cudaMemcpyAsync(cert.data(), ix->g_cert, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(counts.data(), ix->g_count, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0;
    ix->g_ev_pending = false;
    CK(cudaEventElapsedTime(&ms, ev[0], ev[1])); ix->stat_gemm_pass1_us = (int64_t)(ms * 1e3);
    CK(cudaEventElapsedTime(&ms, ev[1], ev[2])); ix->stat_gemm_pass2_us = (int64_t)(ms * 1e3);
    CK(cudaEventElapsedTime(&ms, ev[2], ev[3])); ix->stat_gemm_rerank_us = (int64_t)(ms * 1e3);
    if (depth == 0) {
        CK(cudaEventElapsedTime(&ms, ev[4], ev[0])); ix->stat_gemm_pre_us = (int64_t)(ms * 1e3);
        ix->stat_gemm_host_us = (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - host_t0).count();
    }
    std::vector<int64_t> bad;
    int64_t cand_total = 0;
    for (int64_t i = 0; i < nq; ++i) {
        cand_total += counts[(size_t)i];
        if (!cert[(size_t)i]) bad.push_back(i);
    }
    if (depth == 0) {
        ix->stat_gemm_used = 1;
        ix->stat_gemm_fallbacks = (int64_t)bad.size();
        ix->stat_gemm_cand_total = cand_total;
        ix->stat_gemm_scan_fallbacks = 0;
    } else {
        ix->stat_gemm_scan_fallbacks = (int64_t)bad.size();
    }
    if (!bad.empty()) {
        // gather the failed queries, search them exactly 8 at a time, scatter the results back
        const size_t nb = bad.size();
        DevTmp tq, tD, tI;
        CK(tq.alloc(nb * ix->d * 4));
        CK(tD.alloc(nb * (size_t)k * 4));
        CK(tI.alloc(nb * (size_t)k * 8));
        float* qtmp = (float*)tq.p;
        float* Dtmp = (float*)tD.p;
        int64_t* Itmp = (int64_t*)tI.p;
        for (size_t j = 0; j < nb; ++j)
            CK(cudaMemcpyAsync(qtmp + j * ix->d, q_dev + (size_t)bad[j] * ix->d, (size_t)ix->d * 4, cudaMemcpyDeviceToDevice, st));
        // first retry on the tensor cores with a 3x wider threshold (one more sweep of the shadow serves
        // up to 256 stragglers; the scan kernel needs a pass per 8), then the exact scan for what is left
        const int64_t s_p1 = ix->stat_gemm_pass1_us, s_p2 = ix->stat_gemm_pass2_us, s_rr = ix->stat_gemm_rerank_us;
        int rc = depth == 0 ? search_gemm(ix, qtmp, (int64_t)nb, k, Dtmp, Itmp, st, 1)
                            : search_scan_block(ix, qtmp, (int64_t)nb, k, Dtmp, Itmp, st);
        if (depth == 0) {
            ix->stat_gemm_pass1_us = s_p1;
            ix->stat_gemm_pass2_us = s_p2;
            ix->stat_gemm_rerank_us = s_rr;
        }
        cudaError_t ce = cudaSuccess;
        if (rc == 0)
            for (size_t j = 0; j < nb && ce == cudaSuccess; ++j) {
                ce = cudaMemcpyAsync(D_dev + (size_t)bad[j] * k, Dtmp + j * k, (size_t)k * 4, cudaMemcpyDeviceToDevice, st);
                if (ce == cudaSuccess)
                    ce = cudaMemcpyAsync(I_dev + (size_t)bad[j] * k, Itmp + j * k, (size_t)k * 8, cudaMemcpyDeviceToDevice, st);
            }
        const cudaError_t se = cudaStreamSynchronize(st);  // always: the temporaries are released when this scope ends
        if (rc) return rc;
        CK(ce);
        CK(se);
    }
    return 0;
}


// Single-query pre-filter (option `prefilter`, off by default — the fp32 scan is the path north_star names and the
// one the headline measures).  The SAME scan kernel ranks the resident bf16 shadow (half the bytes of the fp32 rows;
// inner product, for L2 through the two augmented columns carrying |y|^2/2 against -1 in the query) and returns its best
// kp = max(32, 4k) rows; those are re-scored from the fp32 rows with the scan's exact arithmetic (rerank_kernel) and the
// answer is accepted only under the same certificate as the batched path: every row outside the list has approximate
// score <= theta = the kp-th approximate score, hence exact score <= theta + eps with eps = 2^-9 |q||y| (one bf16
// rounding per row element, queries stay fp32) + accumulation slack, so nothing outside can reach the top k when the
// k-th exact score beats that.  Returns 0 = D/I hold the proven exact answer, 2 = not applicable or not certified
// (the caller runs the fp32 scan).  One host read of the 4-byte certificate per search.
static int search_prefilter(b200_index* ix, const float* q_dev, int64_t k, float* D_dev, int64_t* I_dev, cudaStream_t st) {
    const int kp = (int)std::min<int64_t>(B200_FUSED_K_MAX, std::max<int64_t>(32, 4 * k));
    if (4 * k > B200_FUSED_K_MAX || ix->ntotal < 16 * kp || ix->cur_mask || ix->xchg_active || ix->d < 32) return 2;
    if (ix->store != B200_STORE_F32) return 2;  // bf16 rows: the shadow would be the same bytes again
    if (ensure_shadow(ix, st) != 0) return 2;  // resident shadows only
    const int kpad = gemm_kpad(ix);
    const bool l2 = ix->metric == B200_METRIC_L2;
    const int dv = ix->d + (l2 ? 2 : 0);
    // scratch: Dp[kp] f32 | Ip[kp] i64 | cand[kp] u32 | count, theta, qn2, cert | q'[dv]
    const size_t off_I = 256 * 4, off_c = off_I + 256 * 8, off_s = off_c + 256 * 4, off_q = off_s + 64;
    const size_t qx_floats = (size_t)((ix->d + 3) & ~3);
    const size_t need = off_q + (qx_floats + (size_t)dv + 8) * 4;
    if (ix->pf_cap < need) {
        CK(cudaStreamSynchronize(st));
        CKI(grow(&ix->pf_buf, &ix->pf_cap, need));
    }
    if (!ix->pf_cert_host) CK(cudaHostAlloc((void**)&ix->pf_cert_host, 64, cudaHostAllocDefault));
    float* Dp = (float*)ix->pf_buf;
    int64_t* Ip = (int64_t*)(ix->pf_buf + off_I);
    uint32_t* cand = (uint32_t*)(ix->pf_buf + off_c);
    unsigned int* count = (unsigned int*)(ix->pf_buf + off_s);
    float* theta = (float*)(ix->pf_buf + off_s + 16);
    float* qn2 = (float*)(ix->pf_buf + off_s + 32);
    int* cert = (int*)(ix->pf_buf + off_s + 48);
    float* qx = (float*)(ix->pf_buf + off_q);
    const float* q_scan = q_dev;
    if (ix->cur_norm_q) {  // cosine, normalisation deferred to the scan's prologue by the caller: K1 on the query here
                           // (the view below scans with another logical dimension)
        CKI(ingest_dev(ix->d, ix->d, B200_STORE_F32, q_dev, (uint8_t*)qx, (size_t)ix->d * 4, 1, 1, ix->num_sms, st, &ix->launches));
        q_scan = qx;
    }
    const float* q_exact = q_scan;  // what the re-rank scores against (the normalised fp32 query)
    if (l2) {
        float* qe = qx + qx_floats;
        extend_query_kernel<<<(unsigned)((dv + 255) / 256), 256, 0, st>>>(q_scan, 1, ix->d, qe);
        ++ix->launches;
        q_scan = qe;
    }
    {
        // the scan kernel over the shadow: a view with bf16 rows of kpad columns, inner product, row positions as ids
        const ScanView shadow{(const uint8_t*)ix->sh_rows, (size_t)kpad * 2, B200_STORE_BF16, B200_METRIC_IP, dv, kpad,
                              pick_lpr((size_t)kpad * 2 / 16), nullptr, false};
        ScanPlan pl;
        CKI(plan_scan(ix, 1, kp, false, &pl, &shadow));
        CKI(launch_scan(ix, pl, q_scan, 1, kp, Dp, Ip, nullptr, st, &shadow));
    }
    prefilter_lists_kernel<<<1, 256, 0, st>>>(Dp, Ip, kp, q_exact, ix->d, ix->d, cand, count, theta, qn2);
    ++ix->launches;
    RerankParams rp;
    memset(&rp, 0, sizeof rp);
    rp.rows = ix->rows;
    rp.pitch_bytes = ix->pitch;
    rp.nvec = (uint32_t)(ix->pitch / 16);
    rp.store = ix->store;
    rp.d = ix->d;
    rp.qstride = (ix->d_pad + 7) / 8 * 8;
    rp.lpr = ix->lpr;
    rp.q = q_exact;
    rp.cand_rows = cand;
    rp.cand_count = count;
    rp.cand_cap = (uint32_t)kp;
    rp.theta = theta;
    rp.qnorm2 = qn2;
    rp.max_norm2_bits = ix->sh_maxnorm;
    // one bf16 rounding per row element (u = 2^-9 relative, the query stays fp32) + fp32 accumulation slack of both sums
    rp.eps_rel = 0.001953125f * 1.001f + 2.0f * (float)ix->d * 1.2e-7f + 1e-5f;
    rp.n = (uint64_t)ix->ntotal;
    rp.k = (int)k;
    rp.id_map = ix->ids_state == 1 ? ix->ids : nullptr;
    rp.D = D_dev;
    rp.I = I_dev;
    rp.certified = cert;
    const size_t rsmem = (size_t)rp.qstride * 4 + (size_t)next_pow2((uint32_t)std::max(kp, 2)) * 8 + 64;
    if (l2) {
        CK(cudaFuncSetAttribute(rerank_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        rerank_kernel<1><<<1, 256, rsmem, st>>>(rp);
    } else {
        CK(cudaFuncSetAttribute(rerank_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rsmem));
        rerank_kernel<0><<<1, 256, rsmem, st>>>(rp);
    }
    ++ix->launches;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(ix->pf_cert_host, cert, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    ix->stat_prefilter_used = 1;
    if (*ix->pf_cert_host == 1) return 0;
    ++ix->stat_prefilter_fallbacks;
    return 2;
}

static int search_scan_block(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev,
                             cudaStream_t st) {
    int64_t q0 = 0;
    while (q0 < nq) {
        int qb = pick_qb(nq - q0, ix->opt_qb, ix->lpr);
        int nqb = (int)std::min<int64_t>(qb, nq - q0);
        ScanPlan pl;
        CKI(plan_scan(ix, qb, (int)k, false, &pl));
        CKI(launch_scan(ix, pl, q_dev + (size_t)q0 * ix->d, nqb, (int)k, D_dev + (size_t)q0 * k, I_dev + (size_t)q0 * k,
                        nullptr, st));
        q0 += nqb;
    }
    return 0;
}

// exact_only: the scan / full-ranking paths only (what the row-sharded batch protocol falls back to)
static int search_dev_impl(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev, void* stream,
                           bool exact_only);

extern "C" int b200_index_search_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev,
                                     int64_t* I_dev, void* stream) {
    return search_dev_impl(ix, q_dev, nq, k, D_dev, I_dev, stream, false);
}

static int search_dev_impl(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev, int64_t* I_dev, void* stream,
                           bool exact_only) {
    if (!ix) return fail("null index");
    if (nq < 0) return fail("negative nq");
    if (k <= 0) return fail("k must be positive, got %lld", (long long)k);
    if (nq == 0) return 0;
    if (!q_dev || !D_dev || !I_dev) return fail("null buffer");
    if (nq * k > ((int64_t)1 << 40)) return fail("result too large");
    CKI(use_device(ix));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    CKI(order_after_previous_stream(ix, st));
    if (ix->ntotal == 0) {
        int64_t count = nq * k;
        unsigned blocks = (unsigned)((count + 255) / 256);
        if (ix->metric == B200_METRIC_IP)
            fill_pad_kernel<0><<<blocks, 256, 0, st>>>(D_dev, I_dev, count);
        else
            fill_pad_kernel<1><<<blocks, 256, 0, st>>>(D_dev, I_dev, count);
        ++ix->launches;
        CK(cudaGetLastError());
        return 0;
    }
    const bool fullrank = k >= ix->opt_fullrank_min_k || k > B200_FUSED_K_MAX;
    const bool use_gemm = !exact_only && !fullrank && !ix->xchg_active && gemm_eligible(ix, nq, k) && ix->sh_failed_rows != ix->ntotal;
    struct NormGuard {  // the scan launches of THIS search normalise their queries while staging them
        b200_index* ix;
        ~NormGuard() { ix->cur_norm_q = false; }
    } norm_guard{ix};
    if (ix->opt_normalize_queries) {
        if (ix->opt_fuse_query_norm && !use_gemm) {
            ix->cur_norm_q = true;
        } else {  // the tensor-core path builds its bf16 query shadow from normalised queries in memory
            if (ix->qn_cap < (size_t)nq * ix->d) {
                CK(cudaStreamSynchronize(st));
                CKI(grow(&ix->qn_dev, &ix->qn_cap, (size_t)nq * ix->d));
            }
            CKI(ingest_dev(ix->d, ix->d, B200_STORE_F32, q_dev, (uint8_t*)ix->qn_dev, (size_t)ix->d * 4, nq, 1,
                           ix->num_sms, st, &ix->launches));
            q_dev = ix->qn_dev;
        }
    }
    ix->stat_gemm_used = 0;
    ix->stat_prefilter_used = 0;
    if (!exact_only && !fullrank && !use_gemm && nq == 1 && ix->opt_prefilter) {
        const int rc = search_prefilter(ix, q_dev, k, D_dev, I_dev, st);
        if (rc != 2) return rc;
    }
    if (!fullrank) {
        if (use_gemm) {
            // K3 in blocks of at most 16384 queries (bounds the candidate and sample scratch)
            int rc = 0;
            const int64_t qblock = gemm_query_block(ix);
            for (int64_t q0 = 0; q0 < nq && rc == 0; q0 += qblock) {
                int64_t nb = std::min<int64_t>(qblock, nq - q0);
                rc = search_gemm(ix, q_dev + (size_t)q0 * ix->d, nb, k, D_dev + (size_t)q0 * k, I_dev + (size_t)q0 * k, st);
            }
            if (rc != 2) return rc;
            ix->stat_gemm_used = 0;  // no memory for the bf16 shadow: exact scan instead
        }
        return search_scan_block(ix, q_dev, nq, k, D_dev, I_dev, st);
    }
    // full ranking: scores for a block of queries, then one stable radix sort per query
    const uint64_t n = (uint64_t)ix->ntotal;
    int64_t q0 = 0;
    while (q0 < nq) {
        int qb = pick_qb(nq - q0, ix->opt_qb, ix->lpr);
        int nqb = (int)std::min<int64_t>(qb, nq - q0);
        if (ix->fr_hi_cap < (size_t)qb * n) {
            CK(cudaStreamSynchronize(st));
            CKI(grow(&ix->fr_hi, &ix->fr_hi_cap, (size_t)qb * n));
        }
        ScanPlan pl;
        CKI(plan_scan(ix, qb, 1, true, &pl));
        CKI(launch_scan(ix, pl, q_dev + (size_t)q0 * ix->d, nqb, 1, nullptr, nullptr, ix->fr_hi, st));
        if (n <= FULLRANK_SMALL_MAX) {
            const uint32_t m = std::max<uint32_t>(2, next_pow2((uint32_t)n));
            const size_t fsmem = (size_t)m * 8;
            const int64_t* idm = ix->ids_state == 1 ? ix->ids : nullptr;
            if (ix->metric == B200_METRIC_IP) {
                CK(cudaFuncSetAttribute(fullrank_small_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                fullrank_small_kernel<0><<<nqb, 1024, fsmem, st>>>(ix->fr_hi, (uint32_t)n, m, k, idm, D_dev + (size_t)q0 * k,
                                                                   I_dev + (size_t)q0 * k);
            } else {
                CK(cudaFuncSetAttribute(fullrank_small_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem));
                fullrank_small_kernel<1><<<nqb, 1024, fsmem, st>>>(ix->fr_hi, (uint32_t)n, m, k, idm, D_dev + (size_t)q0 * k,
                                                                   I_dev + (size_t)q0 * k);
            }
            ++ix->launches;
            CK(cudaGetLastError());
        } else {
            for (int qi = 0; qi < nqb; ++qi)
                CKI(fullrank_one(ix, ix->fr_hi + (size_t)qi * n, k, D_dev + (size_t)(q0 + qi) * k,
                                 I_dev + (size_t)(q0 + qi) * k, st));
        }
        q0 += nqb;
    }
    return 0;
}

static int search_host_impl(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host, int64_t* I_host, bool exchange);

extern "C" int b200_index_search(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host,
                                 int64_t* I_host) {
    return search_host_impl(ix, q_host, nq, k, D_host, I_host, false);
}

// exchange: the fused multi-GPU form (b200_index_search_exchange_dev) instead of the local search
static int search_host_impl(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host, int64_t* I_host, bool exchange) {
    if (!ix) return fail("null index");
    if (nq < 0) return fail("negative nq");
    if (k <= 0) return fail("k must be positive, got %lld", (long long)k);
    if (nq == 0) return 0;
    if (!q_host || !D_host || !I_host) return fail("null buffer");
    CKI(use_device(ix));
    cudaStream_t st = ix->stream;
    auto search_any = [&](const float* qd, float* Dd, int64_t* Id) -> int {
        return exchange ? b200_index_search_exchange_dev(ix, qd, nq, k, Dd, Id, st) : b200_index_search_dev(ix, qd, nq, k, Dd, Id, st);
    };
    const size_t qn = (size_t)nq * ix->d, on = (size_t)nq * (size_t)k;
    if (ix->q_cap < qn || ix->out_cap < on) CK(cudaStreamSynchronize(st));
    CKI(grow(&ix->q_dev, &ix->q_cap, qn));
    if (ix->out_cap < on) {
        if (ix->D_dev) CK(cudaFree(ix->D_dev));
        if (ix->I_dev) CK(cudaFree(ix->I_dev));
        ix->D_dev = nullptr;
        ix->I_dev = nullptr;
        ix->out_cap = 0;
        CK(cudaMalloc((void**)&ix->D_dev, on * sizeof(float)));
        CK(cudaMalloc((void**)&ix->I_dev, on * sizeof(int64_t)));
        ix->out_cap = on;
    }
    // small transfers go through pinned staging so the copies are truly asynchronous DMA
    const size_t pin_need = qn * 4 + on * 12 + 64;
    const bool use_pin = pin_need <= ((size_t)8 << 20);
    if (use_pin && ix->pin_cap < pin_need) {
        if (ix->pin) CK(cudaFreeHost(ix->pin));
        ix->pin = nullptr;
        ix->pin_cap = 0;
        size_t cap = std::max<size_t>(pin_need, (size_t)1 << 16);
        CK(cudaHostAlloc(&ix->pin, cap, cudaHostAllocDefault));
        ix->pin_cap = cap;
    }
    if (use_pin) {
        uint8_t* base = (uint8_t*)ix->pin;
        int64_t* pI = (int64_t*)base;                       // 8-byte aligned first
        float* pD = (float*)(base + on * 8);
        float* pq = (float*)(base + on * 12);
        memcpy(pq, q_host, qn * 4);
        CK(cudaMemcpyAsync(ix->q_dev, pq, qn * 4, cudaMemcpyHostToDevice, st));
        if (on * 12 <= 4096 && ix->opt_direct_results) {
            // a handful of results (the latency path): the kernels write them straight into the pinned buffer, which is
            // device-addressable under unified addressing — no device->host copies after the search
            CKI(search_any(ix->q_dev, pD, pI));
        } else {
            CKI(search_any(ix->q_dev, ix->D_dev, ix->I_dev));
            CK(cudaMemcpyAsync(pD, ix->D_dev, on * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(pI, ix->I_dev, on * 8, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
        memcpy(D_host, pD, on * 4);
        memcpy(I_host, pI, on * 8);
    } else {
        // big batches / full rankings (memo's k = ntotal): both directions through the pinned ring
        HostSource qs;
        qs.mem = (const uint8_t*)q_host;
        // measured (tools/bench_fullrank.py): 120 MB of results 14.3 ms staged vs 31.4 ms plain; 12 MB 2.9 vs 2.4 ms.
        // (numpy already asks for huge pages on large arrays; advising again changed nothing.)
        const bool staged = ix->opt_staged_results != 0 && on * 12 >= ((size_t)32 << 20);
        const bool staged_q = ix->opt_staged_results != 0;
        if (staged_q && qn * 4 >= ((size_t)4 << 20))
            CKI(upload_staged(ix, qs, qn * 4, (uint8_t*)ix->q_dev, false, 4, [](uint8_t*, size_t, size_t) { return 0; }));
        else
            CK(cudaMemcpyAsync(ix->q_dev, q_host, qn * 4, cudaMemcpyHostToDevice, st));
        CKI(search_any(ix->q_dev, ix->D_dev, ix->I_dev));
        if (staged) {
            CKI(download_staged(ix, (const uint8_t*)ix->D_dev, on * 4, (uint8_t*)D_host, st));
            CKI(download_staged(ix, (const uint8_t*)ix->I_dev, on * 8, (uint8_t*)I_host, st));
        } else {
            CK(cudaMemcpyAsync(D_host, ix->D_dev, on * 4, cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(I_host, ix->I_dev, on * 8, cudaMemcpyDeviceToHost, st));
        }
        CK(cudaStreamSynchronize(st));
    }
    return 0;
}


extern "C" int b200_index_search_masked_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k,
                                            const uint32_t* mask_dev, float* D_dev, int64_t* I_dev, void* stream) {
    if (!ix) return fail("null index");
    ix->cur_mask = mask_dev;
    int rc = b200_index_search_dev(ix, q_dev, nq, k, D_dev, I_dev, stream);
    ix->cur_mask = nullptr;
    return rc;
}

extern "C" int b200_index_search_masked(b200_index* ix, const float* q_host, int64_t nq, int64_t k,
                                        const uint32_t* mask_host, float* D_host, int64_t* I_host) {
    if (!ix) return fail("null index");
    if (!mask_host) return b200_index_search(ix, q_host, nq, k, D_host, I_host);
    CKI(use_device(ix));
    const size_t words = ((size_t)ix->ntotal + 31) / 32;
    if (words == 0) return b200_index_search(ix, q_host, nq, k, D_host, I_host);
    if (ix->mask_cap < words) {
        CK(cudaStreamSynchronize(ix->stream));
        CKI(grow(&ix->mask_dev, &ix->mask_cap, words));
    }
    CK(cudaMemcpyAsync(ix->mask_dev, mask_host, words * 4, cudaMemcpyHostToDevice, ix->stream));
    ix->cur_mask = ix->mask_dev;
    int rc = b200_index_search(ix, q_host, nq, k, D_host, I_host);
    ix->cur_mask = nullptr;
    return rc;
}


extern "C" int b200_index_search_ids_allowed(b200_index* ix, const float* q_host, int64_t nq, int64_t k,
                                             const int64_t* allowed_host, int64_t m, float* D_host, int64_t* I_host) {
    if (!ix) return fail("null index");
    if (m < 0 || (m > 0 && !allowed_host)) return fail("bad allowed-id list");
    if (ix->ntotal == 0) return b200_index_search(ix, q_host, nq, k, D_host, I_host);
    CKI(use_device(ix));
    cudaStream_t st = ix->stream;
    const uint64_t n = (uint64_t)ix->ntotal;
    const size_t words = (size_t)((n + 31) / 32);
    if (ix->mask_cap < words || ix->allow_cap < (size_t)m) CK(cudaStreamSynchronize(st));
    CKI(grow(&ix->mask_dev, &ix->mask_cap, words));
    CKI(grow(&ix->allow_dev, &ix->allow_cap, (size_t)std::max<int64_t>(m, 1)));
    const unsigned blocks = (unsigned)ix->num_sms * 8;
    const bool have_ids = ix->ids_state == 1;
    if (m == 0) {
        CK(cudaMemsetAsync(ix->mask_dev, 0, words * 4, st));
    } else if (!have_ids) {
        // ids are row positions: the row bitmap IS the id bitmap
        CK(cudaMemcpyAsync(ix->allow_dev, allowed_host, (size_t)m * 8, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync(ix->mask_dev, 0, words * 4, st));
        scatter_allowed_kernel<<<blocks, 256, 0, st>>>(ix->allow_dev, (uint64_t)m, 0, n, ix->mask_dev);
        ++ix->launches;
        CK(cudaGetLastError());
    } else {
        if (ix->ids_minmax_rows != ix->ntotal) {  // id range of the map, cached until the next add / reset
            if (!ix->ids_minmax_dev) CK(cudaMalloc((void**)&ix->ids_minmax_dev, 2 * sizeof(long long)));
            const long long init[2] = {LLONG_MAX, LLONG_MIN};
            long long got[2];
            CK(cudaMemcpyAsync(ix->ids_minmax_dev, init, sizeof init, cudaMemcpyHostToDevice, st));
            ids_minmax_kernel<<<blocks, 256, 0, st>>>(ix->ids, n, ix->ids_minmax_dev);
            ++ix->launches;
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(got, ix->ids_minmax_dev, sizeof got, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            ix->ids_min = got[0];
            ix->ids_max = got[1];
            ix->ids_minmax_rows = ix->ntotal;
        }
        const unsigned __int128 span = (unsigned __int128)((__int128)ix->ids_max - (__int128)ix->ids_min) + 1;
        const uint64_t dense_limit = 64 * n + ((uint64_t)1 << 24);  // id bitmap of at most 8 N + 2 MB bytes
        if (span <= dense_limit) {
            const uint64_t range = (uint64_t)span;
            const size_t bwords = (size_t)((range + 31) / 32);
            if (ix->idbits_cap < bwords) CK(cudaStreamSynchronize(st));
            CKI(grow(&ix->idbits_dev, &ix->idbits_cap, bwords));
            CK(cudaMemcpyAsync(ix->allow_dev, allowed_host, (size_t)m * 8, cudaMemcpyHostToDevice, st));
            CK(cudaMemsetAsync(ix->idbits_dev, 0, bwords * 4, st));
            scatter_allowed_kernel<<<blocks, 256, 0, st>>>(ix->allow_dev, (uint64_t)m, ix->ids_min, range, ix->idbits_dev);
            gather_row_mask_kernel<<<blocks, 256, 0, st>>>(ix->ids, n, ix->ids_min, range, ix->idbits_dev, ix->mask_dev);
            ix->launches += 2;
            CK(cudaGetLastError());
        } else {
            // sparse id space: sorted list + one binary search per row
            std::vector<int64_t> sorted(allowed_host, allowed_host + m);
            std::sort(sorted.begin(), sorted.end());
            CK(cudaMemcpyAsync(ix->allow_dev, sorted.data(), (size_t)m * 8, cudaMemcpyHostToDevice, st));
            gather_row_mask_sorted_kernel<<<blocks, 256, 0, st>>>(ix->ids, n, ix->allow_dev, (uint64_t)m, ix->mask_dev);
            ++ix->launches;
            CK(cudaGetLastError());
            CK(cudaStreamSynchronize(st));  // `sorted` is pageable memory that dies with this scope
        }
    }
    ix->cur_mask = ix->mask_dev;
    int rc = b200_index_search(ix, q_host, nq, k, D_host, I_host);
    ix->cur_mask = nullptr;
    return rc;
}


// ---------------------------------------------------------------------------------------------
// fused multi-GPU exchange
// ---------------------------------------------------------------------------------------------
extern "C" size_t b200_exchange_slot_bytes(void) { return XCHG_SLOT_BYTES; }

extern "C" int b200_ipc_alloc(void** out_dev, size_t bytes, char handle_out[64]) {
    if (!out_dev || !handle_out || bytes == 0) return fail("bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail("cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    *out_dev = p;
    return 0;
}
extern "C" int b200_ipc_open(const char handle[64], void** out_dev) {
    if (!handle || !out_dev) return fail("bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CK(cudaIpcOpenMemHandle(out_dev, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}
extern "C" int b200_ipc_close(void* dev) {
    if (dev) CK(cudaIpcCloseMemHandle(dev));
    return 0;
}
extern "C" int b200_ipc_free(void* dev) {
    if (dev) CK(cudaFree(dev));
    return 0;
}
extern "C" int b200_index_set_exchange(b200_index* ix, int world, int rank, void* const* peer_bufs) {
    if (!ix) return fail("null index");
    if (world < 1 || world > 64 || rank < 0 || rank >= world || !peer_bufs) return fail("bad exchange topology");
    CKI(use_device(ix));
    CK(cudaStreamSynchronize(ix->stream));
    if (ix->xchg_peers_dev) CK(cudaFree(ix->xchg_peers_dev));
    ix->xchg_peers_dev = nullptr;
    CK(cudaMalloc((void**)&ix->xchg_peers_dev, (size_t)world * sizeof(void*)));
    CK(cudaMemcpy(ix->xchg_peers_dev, peer_bufs, (size_t)world * sizeof(void*), cudaMemcpyHostToDevice));
    if (!ix->xchg_status_host) {
        int* h = nullptr;
        CK(cudaHostAlloc((void**)&h, sizeof(int), cudaHostAllocMapped));
        *h = 0;
        CK(cudaHostGetDevicePointer((void**)&ix->xchg_status, h, 0));
        ix->xchg_status_host = h;
    }
    ix->xchg_world = world;
    ix->xchg_rank = rank;
    ix->xchg_epoch = 0;
    return 0;
}
extern "C" int b200_index_search_exchange_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev,
                                              int64_t* I_dev, void* stream) {
    if (!ix) return fail("null index");
    if (!ix->xchg_peers_dev) return fail("b200_index_set_exchange has not been called");
    if (k > B200_FUSED_K_MAX || k >= ix->opt_fullrank_min_k) return fail("fused exchange needs k <= %d", B200_FUSED_K_MAX);
    if (ix->ntotal == 0) return fail("fused exchange needs at least one row on every rank");
    if (*ix->xchg_status_host) return fail("fused exchange: a peer GPU did not deliver its results in time during an earlier search");
    ix->xchg_active = true;
    int rc = b200_index_search_dev(ix, q_dev, nq, k, D_dev, I_dev, stream);
    ix->xchg_active = false;
    return rc;
}

// 0 = every fused exchange so far completed; 1 = a peer did not deliver in time (the affected search returned
// padding only).  Read it after synchronising the stream the search ran on.
// host query -> host result through the fused exchange, on the handle's own stream (one call = staging copy, H2D, the
// kernel, results written straight into pinned host memory, one synchronisation, the exchange status check)
extern "C" int b200_index_search_exchange(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host, int64_t* I_host) {
    int rc = search_host_impl(ix, q_host, nq, k, D_host, I_host, true);
    if (rc) return rc;
    if (ix->xchg_status_host && *ix->xchg_status_host)
        return fail("fused exchange: a peer GPU did not deliver its results in time; the search returned no results");
    return 0;
}

extern "C" int b200_index_exchange_status(b200_index* ix) {
    if (!ix) return -1;
    return ix->xchg_status_host ? *ix->xchg_status_host : 0;
}

// Phase stamps (globaltimer, ns) of the last scan launch made with option scan_phase_stamps = 1:
// out[cta * 8 + j], j = 0 kernel entry, 1 queries staged, 2 scan done (warp 0), 3 CTA reduction written, 4 last CTA
// starts the final merge, 5 final merge done, 6 kernel end (last CTA), 7 first tile landed (warp 0); 0 = not reached.
extern "C" int b200_index_read_phase_stamps(b200_index* ix, unsigned long long* out_host, int64_t cap_words, int64_t* n_ctas) {
    if (!ix || !out_host || !n_ctas) return fail("null argument");
    *n_ctas = 0;
    if (!ix->stamps || ix->stamps_grid == 0) return fail("no phase stamps recorded (set option scan_phase_stamps)");
    if (cap_words < (int64_t)ix->stamps_grid * 8) return fail("buffer too small for %d CTAs", ix->stamps_grid);
    CKI(use_device(ix));
    cudaStream_t st = ix->last_stream ? ix->last_stream : ix->stream;
    CK(cudaMemcpyAsync(out_host, ix->stamps, (size_t)ix->stamps_grid * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    *n_ctas = ix->stamps_grid;
    return 0;
}

extern "C" int64_t b200_index_launch_count(b200_index* ix) { return ix ? ix->launches : -1; }
extern "C" int b200_index_sync(b200_index* ix) {
    if (!ix) return fail("null index");
    CKI(use_device(ix));
    CKI(order_after_previous_stream(ix, ix->stream));
    CK(cudaStreamSynchronize(ix->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
// introspection
// ---------------------------------------------------------------------------------------------
extern "C" int64_t b200_index_ntotal(b200_index* ix) { return ix ? ix->ntotal : -1; }
extern "C" int b200_index_d(b200_index* ix) { return ix ? ix->d : -1; }
extern "C" int b200_index_metric(b200_index* ix) { return ix ? ix->metric : -1; }
extern "C" int b200_index_store(b200_index* ix) { return ix ? ix->store : -1; }
extern "C" int b200_index_has_ids(b200_index* ix) { return ix ? (ix->ids_state == 1) : -1; }

extern "C" int b200_index_get_ids(b200_index* ix, int64_t* out_host) {
    if (!ix) return fail("null index");
    if (ix->ntotal == 0) return 0;
    if (!out_host) return fail("null buffer");
    CKI(use_device(ix));
    if (ix->ids_state == 1) {
        CK(cudaMemcpyAsync(out_host, ix->ids, (size_t)ix->ntotal * sizeof(int64_t), cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
    } else {
        for (int64_t i = 0; i < ix->ntotal; ++i) out_host[i] = i;
    }
    return 0;
}

extern "C" int b200_index_get_rows(b200_index* ix, int64_t row0, int64_t n, float* out_host) {
    if (!ix) return fail("null index");
    if (row0 < 0 || n < 0 || row0 + n > ix->ntotal) return fail("row range [%lld,+%lld) out of bounds", (long long)row0, (long long)n);
    if (n == 0) return 0;
    if (!out_host) return fail("null buffer");
    CKI(use_device(ix));
    const uint8_t* src = ix->rows + (size_t)row0 * ix->pitch;
    if (ix->store == B200_STORE_F32) {
        CK(cudaMemcpy2DAsync(out_host, (size_t)ix->d * 4, src, ix->pitch, (size_t)ix->d * 4, (size_t)n,
                             cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
    } else {
        std::vector<uint16_t> tmp((size_t)n * ix->d_pad);
        CK(cudaMemcpyAsync(tmp.data(), src, (size_t)n * ix->pitch, cudaMemcpyDeviceToHost, ix->stream));
        CK(cudaStreamSynchronize(ix->stream));
        for (int64_t r = 0; r < n; ++r)
            for (int c = 0; c < ix->d; ++c) {
                uint32_t u = (uint32_t)tmp[(size_t)r * ix->d_pad + c] << 16;
                memcpy(&out_host[(size_t)r * ix->d + c], &u, 4);
            }
    }
    return 0;
}

extern "C" int b200_index_rows_dev(b200_index* ix, void** out_ptr, size_t* out_pitch_bytes) {
    if (!ix || !out_ptr || !out_pitch_bytes) return fail("null argument");
    *out_ptr = ix->rows;
    *out_pitch_bytes = ix->pitch;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// stand-alone stages
// ---------------------------------------------------------------------------------------------
extern "C" int b200_normalize_rows(float* x_host, int64_t n, int d, int device) {
    if (n < 0 || d <= 0) return fail("bad shape");
    if (n == 0) return 0;
    if (!x_host) return fail("null buffer");
    int count = 0;
    CK(cudaGetDeviceCount(&count));
    if (count <= 0) return fail("no CUDA device: this library has no CPU fallback");
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int d_pad = (d + 3) / 4 * 4;
    float *src = nullptr, *dst = nullptr;
    CK(cudaMalloc((void**)&src, (size_t)n * d * 4));
    cudaError_t e = cudaMalloc((void**)&dst, (size_t)n * d_pad * 4);
    if (e != cudaSuccess) {
        cudaFree(src);
        return fail("cudaMalloc: %s", cudaGetErrorString(e));
    }
    int rc = 0;
    do {
        if ((e = cudaMemcpy(src, x_host, (size_t)n * d * 4, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        rc = ingest_dev(d, d_pad, B200_STORE_F32, src, (uint8_t*)dst, (size_t)d_pad * 4, n, 1, sms, 0, nullptr);
        if (rc) break;
        e = cudaMemcpy2D(x_host, (size_t)d * 4, dst, (size_t)d_pad * 4, (size_t)d * 4, (size_t)n, cudaMemcpyDeviceToHost);
    } while (0);
    cudaFree(src);
    cudaFree(dst);
    if (e != cudaSuccess) return fail("normalize_rows: %s", cudaGetErrorString(e));
    return rc;
}

// Row-sharded batch, local half (sharded.py): this index is one of `world` row shards.  Leaves in D/I the exact
// fp32 scores of this shard's best candidates (best-first, padded) and in bound_dev[nq] the score no row outside
// the list can beat; everything is enqueued on `stream`, nothing is read back.  Shards where the tensor-core path
// does not apply (few rows, k > 256, no room for the shadow, a filter) answer with their exact top k from the scan
// kernel and a bound that excludes nothing.  widen: 0 = first attempt, 1 = second attempt for queries the merged
// certificate rejected (3x more candidates), 2 = exact scan only (last resort; bound excludes nothing).
extern "C" int b200_index_search_shard_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, int world, int widen,
                                           float* D_dev, int64_t* I_dev, float* bound_dev, void* stream) {
    if (!ix) return fail("null index");
    if (nq < 0) return fail("negative nq");
    if (k <= 0) return fail("k must be positive, got %lld", (long long)k);
    if (world < 1) return fail("world must be >= 1");
    if (nq == 0) return 0;
    if (!q_dev || !D_dev || !I_dev || !bound_dev) return fail("null buffer");
    CKI(use_device(ix));
    cudaStream_t st = stream ? (cudaStream_t)stream : ix->stream;
    const bool fullrank = k >= ix->opt_fullrank_min_k || k > B200_FUSED_K_MAX;
    // at least 2 queries' worth of tensor work even when gemm_min_nq is lower; the caller decides what a batch is
    const bool use_gemm = widen != 2 && !fullrank && !ix->xchg_active && ix->ntotal > 0 &&
                          gemm_eligible(ix, std::max<int64_t>(nq, ix->opt_gemm_min_nq), k) && ix->sh_failed_rows != ix->ntotal &&
                          ix->cur_mask == nullptr;
    if (use_gemm) {
        CKI(order_after_previous_stream(ix, st));
        const float* q_raw = q_dev;
        if (ix->opt_normalize_queries) {  // cosine: the bf16 query shadow is built from normalised queries (K1)
            if (ix->qn_cap < (size_t)nq * ix->d) {
                CK(cudaStreamSynchronize(st));
                CKI(grow(&ix->qn_dev, &ix->qn_cap, (size_t)nq * ix->d));
            }
            CKI(ingest_dev(ix->d, ix->d, B200_STORE_F32, q_dev, (uint8_t*)ix->qn_dev, (size_t)ix->d * 4, nq, 1, ix->num_sms, st,
                           &ix->launches));
            q_dev = ix->qn_dev;
        }
        int rc = 0;
        const int64_t qblock = gemm_query_block(ix);
        for (int64_t q0 = 0; q0 < nq && rc == 0; q0 += qblock) {
            const int64_t nb = std::min<int64_t>(qblock, nq - q0);
            rc = search_gemm(ix, q_dev + (size_t)q0 * ix->d, nb, k, D_dev + (size_t)q0 * k, I_dev + (size_t)q0 * k, st,
                             widen == 1 ? 1 : 0, bound_dev + q0, world);
            if (rc == 2 && q0 != 0) return fail("the bf16 shadow disappeared between two blocks of one batch");
        }
        if (rc != 2) return rc;
        q_dev = q_raw;
    }
    // exact local answer + neutral bound
    const unsigned blocks = (unsigned)((nq + 255) / 256);
    if (ix->metric == B200_METRIC_IP) fill_neutral_bound_kernel<0><<<blocks, 256, 0, st>>>(bound_dev, nq);
    else fill_neutral_bound_kernel<1><<<blocks, 256, 0, st>>>(bound_dev, nq);
    CK(cudaGetLastError());
    return search_dev_impl(ix, q_dev, nq, k, D_dev, I_dev, stream, /*exact_only=*/true);
}

// K4 + certificate of a row-sharded batch: merges the G shard lists like b200_merge_topk_dev, then marks the queries
// whose merged k-th entry does not strictly beat every shard's bound (bounds_parts_dev: shard g's bounds at
// bounds_parts_dev + g * bound_part_stride floats).  n_total = rows in all shards.  uncertified_dev[nq] gets 0/1,
// *n_uncertified_dev the count (zeroed here).
extern "C" int b200_merge_certify_dev(int metric, int G, int64_t nq, int64_t k, int64_t n_total, const float* D_parts_dev,
                                      const int64_t* I_parts_dev, int64_t D_part_stride, int64_t I_part_stride,
                                      const float* bounds_parts_dev, int64_t bound_part_stride, float* D_out_dev,
                                      int64_t* I_out_dev, int* uncertified_dev, int* n_uncertified_dev, void* stream) {
    if (!bounds_parts_dev || !uncertified_dev || !n_uncertified_dev) return fail("null buffer");
    if (n_total < 0) return fail("negative n_total");
    int rc = b200_merge_topk_dev(metric, G, nq, k, D_parts_dev, I_parts_dev, D_part_stride, I_part_stride, D_out_dev, I_out_dev, stream);
    if (rc || nq == 0) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    CK(cudaMemsetAsync(n_uncertified_dev, 0, sizeof(int), st));
    const int64_t want = std::min<int64_t>(k, n_total);
    const int64_t bs = bound_part_stride ? bound_part_stride : nq;
    const unsigned blocks = (unsigned)((nq + 255) / 256);
    if (metric == B200_METRIC_IP)
        merge_certify_kernel<0><<<blocks, 256, 0, st>>>(G, nq, k, want, D_out_dev, bounds_parts_dev, bs, uncertified_dev, n_uncertified_dev);
    else
        merge_certify_kernel<1><<<blocks, 256, 0, st>>>(G, nq, k, want, D_out_dev, bounds_parts_dev, bs, uncertified_dev, n_uncertified_dev);
    CK(cudaGetLastError());
    return 0;
}

extern "C" int b200_merge_topk_dev(int metric, int G, int64_t nq, int64_t k, const float* D_parts_dev,
                                   const int64_t* I_parts_dev, int64_t D_part_stride, int64_t I_part_stride,
                                   float* D_out_dev, int64_t* I_out_dev, void* stream) {
    if (G <= 0 || nq < 0 || k <= 0) return fail("bad shape");
    if (nq == 0) return 0;
    if (!D_parts_dev || !I_parts_dev || !D_out_dev || !I_out_dev) return fail("null buffer");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ds = D_part_stride ? D_part_stride : nq * k, is = I_part_stride ? I_part_stride : nq * k;
    int64_t total = (int64_t)G * nq * k;
    unsigned blocks = (unsigned)std::min<int64_t>((total + 255) / 256, 148 * 16);
    if (metric == B200_METRIC_IP)
        merge_topk_kernel<0><<<blocks, 256, 0, st>>>(G, nq, k, D_parts_dev, I_parts_dev, ds, is, D_out_dev, I_out_dev);
    else if (metric == B200_METRIC_L2)
        merge_topk_kernel<1><<<blocks, 256, 0, st>>>(G, nq, k, D_parts_dev, I_parts_dev, ds, is, D_out_dev, I_out_dev);
    else
        return fail("unknown metric %d", metric);
    CK(cudaGetLastError());
    return 0;
}
