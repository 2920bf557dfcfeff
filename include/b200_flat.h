/*
 * b200_flat.h — C ABI of the B200-native flat vector recall path.
 *
 * This is the drop-in boundary for the one hot path of memo (mikesmullin/c99-vectordb v2):
 * the flat (exhaustive) IP / L2 / cosine search and the index rebuild (add) that
 * /root/reference/memo_cli.py drives through the `faiss` module object it imports at
 * memo_cli.py:13.  Every entry point below names the reference interface it replaces.
 *
 * Conventions
 *   - plain C types only: pointers, sizes, ints.  No C++ exceptions cross this boundary.
 *   - every function returns an int status: 0 = ok, non-zero = failure; the message for the
 *     calling thread is then available from b200_last_error().
 *   - "host" pointers are ordinary process memory (numpy buffers); "dev" pointers are CUDA
 *     device pointers on the index's device (torch tensors' data_ptr()).
 *   - one CUDA stream per handle; calls on one handle are serialised by the caller
 *     (memo is single-threaded, memo_cli.py:883-949).  Different handles are independent:
 *     there is no global mutable state.
 *   - there is NO CPU fallback: with no usable CUDA device every compute entry fails loudly.
 *
 * Tie rule (stated): results are ordered by (score best-first, then smaller
 * row position first); at the k-th boundary the smaller row position is kept.
 */
#ifndef B200_FLAT_H
#define B200_FLAT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 1

/* metric: faiss MetricType values [upstream faiss/MetricType.h]: IP = 0, L2 = 1 */
#define B200_METRIC_IP 0
#define B200_METRIC_L2 1

/* row storage */
#define B200_STORE_F32 0
#define B200_STORE_BF16 1

/* scan kernel selection (b200_index_set_option "scan_variant") */
#define B200_SCAN_AUTO 0
#define B200_SCAN_BULK 1 /* cp.async.bulk (TMA) staged smem ring */
#define B200_SCAN_LDG 2  /* direct 128-bit ld.global.nc */

typedef struct b200_index b200_index;

/* ---- library ---------------------------------------------------------------------------- */

int b200_abi_version(void);
/* message of the last failure on the calling thread ("" if none) */
const char* b200_last_error(void);
/* number of CUDA devices visible; fails (non-zero) when the CUDA runtime is unusable */
int b200_device_count(int* out_count);

/* ---- index lifetime ---------------------------------------------------------------------
 * replaces: faiss.IndexHNSWFlat(DIM, 32) / faiss.IndexIDMap2(base) construction,
 *           memo_cli.py:244-248 (restated as IndexFlatIP / IndexFlatL2 per north_star). */
int b200_index_create(b200_index** out, int d, int metric, int store, int device);
int b200_index_destroy(b200_index* ix);
/* drop all rows, keep the allocation (faiss Index::reset [upstream]) */
int b200_index_reset(b200_index* ix);
/* make room for n_total rows without reallocation */
int b200_index_reserve(b200_index* ix, int64_t n_total);
/* Named integer options (tuning knobs for the sweep harness; defaults are the measured best):
 *   scan_variant          0 auto | 1 TMA-staged ring | 2 direct 128-bit loads
 *   scan_warps, scan_stages, scan_tile_rows, scan_ctas_per_sm, scan_l2_evict_first,
 *   scan_query_block (1|2|4|8), scan_dynamic_tiles (-1 auto|0|1), scan_claim_chunk, scan_fused_tail
 *   fullrank_min_k        k at or above which the full-ranking (radix sort) path is used (default 257)
 *   normalize_queries     1: L2-normalise queries on the device before searching (cosine)
 *   host_staged_results   1 (default): results of 32 MB or more (memo's k = ntotal) and query blocks of 4 MB or more
 *                         cross PCIe through the pinned ring with threaded host copies; 0: plain copies
 *   gemm_min_nq           batched tensor-core path (K3) for nq >= this (default 2; 0 disables)
 *   gemm_min_rows, gemm_emit_factor, gemm_sample_tiles, gemm_chunk_tiles, gemm_cta_group (1|2)
 *   gemm_shadow_max_rows  > 0: keep at most this many rows of K3's bf16 shadow resident and stream the rest through it
 *                         chunk by chunk (what happens by itself when the shadow does not fit next to the rows)
 *   prefilter             1: single queries (k <= 64, unfiltered) rank the resident bf16 shadow with the scan kernel
 *                         first — half the bytes of the fp32 rows — re-score the best max(32, 4k) rows from the fp32
 *                         rows and accept the answer only when a certificate proves it exact; otherwise the fp32 scan
 *                         runs.  Same ids and distances as the default (0) either way; costs the shadow (+50 % memory)
 *                         and one 4-byte host read per search
 *   scan_pdl, queries_stable   programmatic dependent launch of back-to-back searches (see b200_index_search_dev)
 * Read-only statistics of the last search (b200_index_get_option): stat_gemm_used, stat_gemm_streamed,
 *   stat_gemm_fallbacks (queries retried), stat_gemm_scan_fallbacks (queries recomputed by the scan),
 *   stat_gemm_cand_total, stat_gemm_pass1_us, stat_gemm_pass2_us, stat_gemm_rerank_us, stat_prefilter_used,
 *   stat_prefilter_fallbacks (searches the certificate sent to the fp32 scan, cumulative). */
int b200_index_set_option(b200_index* ix, const char* name, int64_t value);
int b200_index_get_option(b200_index* ix, const char* name, int64_t* out_value);

/* ---- add (index rebuild) ------------------------------------------------------------------
 * replaces: index.add_with_ids(x[n,d] float32, ids[n] int64), memo_cli.py:282 and :437,
 *           and (normalize=1) memo's normalize(), memo_cli.py:131-135, run at add time.
 * x is row-major contiguous [n,d]; ids may be NULL (then id == row position, faiss Index::add).
 * Both arrays are copied; the caller keeps ownership.  Mixing NULL and non-NULL ids on one
 * index is an error once rows exist. */
int b200_index_add(b200_index* ix, const float* x_host, int64_t n, const int64_t* ids_host,
                   int normalize);
int b200_index_add_dev(b200_index* ix, const float* x_dev, int64_t n, const int64_t* ids_dev,
                       int normalize);
/* ---- .memo payload I/O (fast load / save) -------------------------------------------------
 * replaces: the bulk byte movement inside faiss.read_index(path), memo_cli.py:255, and
 *           faiss.write_index(index, path), memo_cli.py:361 and :448.  The host side parses / writes
 *           the small faiss headers; these move the payload without intermediate host copies:
 * add_file:   n dense float32 rows at byte rows_offset of `path` (and n int64 ids at ids_offset, or
 *             ids_offset < 0 for none) go file -> pinned ring (threaded pread) -> device (-> K1).
 *             A file shorter than rows_offset + n*d*4 (or the id range) is an error, nothing is added.
 * write_file: rows [0, ntotal) as dense float32 at rows_offset (bf16 rows widened, lossless) and the
 *             ids at ids_offset (>= 0) of an EXISTING file: device -> pinned ring -> threaded pwrite. */
int b200_index_add_file(b200_index* ix, const char* path, int64_t rows_offset, int64_t n,
                        int64_t ids_offset, int normalize);
int b200_index_write_file(b200_index* ix, const char* path, int64_t rows_offset, int64_t ids_offset);

/* ---- whole .memo files (faiss index serialisation) -------------------------
 * replaces: faiss.read_index(str) memo_cli.py:255 / faiss.write_index(index, str) memo_cli.py:361, :448
 *           (faiss's own C API: faiss_read_index_fname / faiss_write_index_fname [upstream]).
 * Layouts understood: IndexIDMap ("IxMp") / IndexIDMap2 ("IxM2") over a flat index ("IxFI" / "IxF2" /
 * "IxFl"), a bare flat index, and memo's original files whose nested index is IndexHNSWFlat ("IHNf"):
 * the graph is skipped and the flat storage behind it is used (every size is checked).  Written files
 * are plain flat indexes ("IxFI" for IP, "IxF2" for L2) with or without the id-map wrapper.
 *   b200_memo_probe          parse the headers only (no device needed): what is in the file and where
 *   b200_memo_write_headers  create/truncate `path`, write every header (and the id count), report the
 *                            payload offsets that b200_index_write_file then fills
 *   b200_index_save          headers + payload of a resident index; kind 0 = bare flat, 1 = IxMp, 2 = IxM2
 *                            (an index added without ids saves its row positions as ids)
 *   b200_index_load          probe + create + add_file: a new handle holding the file's rows (and ids) */
typedef struct b200_memo_info {
    int32_t kind;        /* 0 = flat index, 1 = IndexIDMap, 2 = IndexIDMap2 */
    int32_t d;
    int32_t metric;      /* B200_METRIC_* of the flat payload */
    int32_t from_hnsw;   /* 1: the payload sat behind an IndexHNSWFlat graph that was skipped */
    int64_t ntotal;
    int64_t rows_offset; /* byte offset of ntotal * d float32 */
    int64_t ids_offset;  /* byte offset of ntotal int64, or -1 */
} b200_memo_info;
int b200_memo_probe(const char* path, b200_memo_info* out);
int b200_memo_write_headers(const char* path, const b200_memo_info* in, int64_t* rows_offset, int64_t* ids_offset);
int b200_index_save(b200_index* ix, const char* path, int kind);
int b200_index_load(b200_index** out, const char* path, int store, int device, b200_memo_info* info_or_null);
/* synthetic rows generated on the device by the counter-based generator of DESIGN.md §6
 * (bit-identical to oracle/flat_oracle.c:oracle_synth_rows); rows get ids first_id + i when
 * with_ids != 0.  Used by bench.py / tests for databases too large to upload. */
int b200_index_add_synthetic(b200_index* ix, int64_t n, uint64_t seed, int64_t first_row,
                             int normalize, int with_ids, int64_t first_id);

/* ---- search -------------------------------------------------------------------------------
 * replaces: D, I = index.search(x[nq,d] float32, k), memo_cli.py:292
 *           (IndexIDMap2::search over IndexFlatIP/IndexFlatL2 [upstream]).
 * D is float32 [nq,k], I is int64 [nq,k], best-first, padded with id -1 and
 * -FLT_MAX (IP) / +FLT_MAX (L2).  Any k >= 1 is accepted (memo asks for k = ntotal,
 * memo_cli.py:291).  Results are exact on every path: single queries use the HBM-bound scan kernel,
 * batches the tcgen05 tensor-core path with an exact fp32 re-rank and a certificate (uncertified
 * queries are recomputed by the scan), k > 256 the full-ranking radix sort. */
int b200_index_search(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host,
                      int64_t* I_host);
/* device-resident variant: q, D, I are device pointers; work is enqueued on `stream`
 * (a cudaStream_t; NULL = the handle's own stream).  Single queries (nq < option gemm_min_nq), masked
 * scans and k > 256 are enqueued and NOT synchronised.  Batches on the tensor-core path block the host
 * once per call: the certificates are read back to pick the (rare) queries that are recomputed.
 * A handle owns one set of scratch buffers: calls are serialised by the caller on the host, and when
 * consecutive calls name different streams the library orders the new stream behind the previous one
 * (event record + wait), so streams passed here must stay alive while the handle may still use them.
 * Back-to-back searches on one stream overlap through programmatic dependent launch (option scan_pdl,
 * default on): the next launch ramps up while the previous one merges; with option queries_stable = 1
 * (a promise that queries are never produced by the kernel immediately preceding the search on its
 * stream) the whole scan overlaps. */
int b200_index_search_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, float* D_dev,
                          int64_t* I_dev, void* stream);
/* filtered search (push memo's metadata filter down into the scan instead of
 * k = ntotal + Python post-filter, memo_cli.py:291, :491-521).  mask is a bitmap over ROW
 * POSITIONS, ceil(ntotal/32) uint32 words, bit (r & 31) of word r >> 5 set = row r may be returned;
 * NULL = no filter.  Single queries use the scan kernel, batches the tensor-core path (the bitmap is
 * applied in both epilogues); results are exact either way. */
int b200_index_search_masked(b200_index* ix, const float* q_host, int64_t nq, int64_t k,
                             const uint32_t* mask_host, float* D_host, int64_t* I_host);
int b200_index_search_masked_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k,
                                 const uint32_t* mask_dev, float* D_dev, int64_t* I_dev, void* stream);
/* the same filter given as a list of m allowed RECORD IDS (any order, duplicates and unknown ids are
 * fine; m = 0 allows nothing): the row bitmap is built on the device — ids scattered into a bitmap over
 * the id range, one lookup per row (dense id spaces such as memo's record positions, memo_cli.py:276),
 * or a sorted list + binary search per row (sparse ids) — and the masked search runs on it.  Replaces
 * the O(ntotal) Python post-filter loop of memo_cli.py:491-521 without any per-row host work. */
int b200_index_search_ids_allowed(b200_index* ix, const float* q_host, int64_t nq, int64_t k,
                                  const int64_t* allowed_host, int64_t m, float* D_host, int64_t* I_host);
/* ---- fused multi-GPU exchange (one process per GPU) --------------------------------------------
 * Instead of an NCCL all-gather + merge kernel, the scan kernel's last CTA stores its local top-k
 * into every rank's exchange buffer over NVLink peer mappings, flags it, waits for the peers and
 * merges: scan + exchange + merge are ONE kernel per GPU.
 *   b200_ipc_alloc   device buffer (zeroed) + its 64-byte CUDA IPC handle, to be shared with peers
 *   b200_ipc_open    map a peer's handle into this process (enables peer access lazily)
 *   b200_index_set_exchange   peer_bufs[g] = rank g's buffer as mapped here (own buffer for g=rank);
 *                    every buffer holds 2 * world * b200_exchange_slot_bytes() bytes
 *   b200_index_search_exchange_dev   like search_dev for nq <= 8 per launch group and k <= 256, but D/I
 *                    receive the GLOBAL merged result; every rank must call it for every search
 *   b200_index_exchange_status       0 = every exchange so far completed; 1 = a peer did not deliver within
 *                    ~2 s: that search returned padding only (ids -1).  Read after synchronising the stream. */
int b200_ipc_alloc(void** out_dev, size_t bytes, char handle_out[64]);
int b200_ipc_open(const char handle[64], void** out_dev);
int b200_ipc_close(void* dev);
int b200_ipc_free(void* dev);
size_t b200_exchange_slot_bytes(void);
int b200_index_set_exchange(b200_index* ix, int world, int rank, void* const* peer_bufs);
int b200_index_search_exchange_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k,
                                   float* D_dev, int64_t* I_dev, void* stream);
/* host query -> host result through the fused exchange on the handle's own stream (collective: every rank calls it
 * with the same query); fails when a peer did not deliver in time */
int b200_index_search_exchange(b200_index* ix, const float* q_host, int64_t nq, int64_t k, float* D_host, int64_t* I_host);
int b200_index_exchange_status(b200_index* ix);
/* profiling aid: with option scan_phase_stamps = 1 every scan launch records per-CTA globaltimer stamps
 * (ns): out[cta*8 + j], j = 0 kernel entry, 1 queries staged, 2 scan done (warp 0), 3 CTA reduction
 * written, 4 last CTA starts the final merge, 5 final merge done, 6 kernel end (last CTA), 7 first tile
 * landed (warp 0); 0 = not reached.  n_ctas receives the grid size of the last scan launch. */
int b200_index_read_phase_stamps(b200_index* ix, unsigned long long* out_host, int64_t cap_words, int64_t* n_ctas);
/* kernel launches issued by this handle since creation (bench.py's gpu_launches) */
int64_t b200_index_launch_count(b200_index* ix);
/* block until the handle's stream is idle */
int b200_index_sync(b200_index* ix);

/* ---- introspection ------------------------------------------------------------------------
 * replaces: index.ntotal (memo_cli.py:266,:289,:291,:473);
 *           faiss.vector_to_array(index.id_map) (memo_cli.py:268);
 *           the row payload faiss.write_index serialises (memo_cli.py:361,:448). */
int64_t b200_index_ntotal(b200_index* ix);
int b200_index_d(b200_index* ix);
int b200_index_metric(b200_index* ix);
int b200_index_store(b200_index* ix);
int b200_index_has_ids(b200_index* ix);
int b200_index_get_ids(b200_index* ix, int64_t* out_host /* [ntotal] */);
/* rows [row0, row0+n) as float32 [n,d] (bf16 storage is widened exactly) */
int b200_index_get_rows(b200_index* ix, int64_t row0, int64_t n, float* out_host);
/* device pointer to the row storage and its pitch in bytes (for zero-copy callers) */
int b200_index_rows_dev(b200_index* ix, void** out_ptr, size_t* out_pitch_bytes);

/* ---- stand-alone stages ---------------------------------------------------------------------
 * K1 as a function: in-place L2 normalisation of host rows on the device
 * (faiss.normalize_L2 surface; memo normalize(), memo_cli.py:131-135: norm <= 1e-8 -> zeros). */
int b200_normalize_rows(float* x_host, int64_t n, int d, int device);
/* K4/K5: merge G per-shard result lists (shard-major [G,nq,k], each list best-first) into
 * [nq,k]; ties go to the lower shard, then to the earlier position in that shard's list, which
 * is the global smaller-row-first rule for contiguous row-range shards.  Device pointers.
 * Shard g's lists start at D_parts + g*D_part_stride / I_parts + g*I_part_stride (in elements;
 * 0 means the dense nq*k), so one packed all-gather buffer can be merged in place. */
int b200_merge_topk_dev(int metric, int G, int64_t nq, int64_t k, const float* D_parts_dev,
                        const int64_t* I_parts_dev, int64_t D_part_stride, int64_t I_part_stride,
                        float* D_out_dev, int64_t* I_out_dev, void* stream);
/* Row-sharded batches on the tensor-core path (sharded.py; no analogue in memo_cli.py, which is one process):
 * b200_index_search_shard_dev is the local half on one of `world` row shards — D/I receive the exact fp32 scores
 * of this shard's best candidates (best-first, padded like b200_index_search) and bound_dev[nq] the score that no
 * row outside the list can beat; thresholds aim at 1/world of the candidates a single index would collect, so the
 * exact re-rank shrinks with the shard.  Everything is enqueued on `stream`; nothing is read back.  Shards the
 * tensor-core path does not serve (few rows, k > 256, no room for the bf16 shadow) answer with their exact top k
 * and a bound that excludes nothing.  widen: 0 = first attempt, 1 = second attempt with 3x more candidates, 2 = exact scan
 * only (the last resort for queries whose certificate failed twice).
 * b200_merge_certify_dev merges the gathered shard lists (as b200_merge_topk_dev) and takes the certificate over
 * all shards: uncertified_dev[q] = 1 unless the merged k-th entry (k clipped to n_total) strictly beats every
 * shard's bound; *n_uncertified_dev counts them.  Uncertified queries are searched again by the caller (widened,
 * then with the exact scan), so results stay exact. */
int b200_index_search_shard_dev(b200_index* ix, const float* q_dev, int64_t nq, int64_t k, int world, int widen,
                                float* D_dev, int64_t* I_dev, float* bound_dev, void* stream);
int b200_merge_certify_dev(int metric, int G, int64_t nq, int64_t k, int64_t n_total, const float* D_parts_dev,
                           const int64_t* I_parts_dev, int64_t D_part_stride, int64_t I_part_stride,
                           const float* bounds_parts_dev, int64_t bound_part_stride, float* D_out_dev,
                           int64_t* I_out_dev, int* uncertified_dev, int* n_uncertified_dev, void* stream);
/* Host-side bulk hashing-trick embedder (replaces the token loop of embed_text_hash,
 * memo_cli.py:158-166) with CPython's str hash fixed to the PYTHONHASHSEED=0 key: n lower-cased
 * UTF-8 texts concatenated in utf8, text i = [offsets[i], offsets[i+1]); out is [n,dim] float32
 * un-normalised bucket counts.  b200_py_hash_seed0 is the token hash itself. */
int b200_hash_embed(const char* utf8, const int64_t* offsets, int64_t n, int dim, float* out);
int64_t b200_py_hash_seed0(const char* bytes, int64_t len);
/* K6 — bulk rebuild from texts on the device (replaces rebuild_index_from_texts, memo_cli.py:272-285, and the
 * embed_text_hash + normalize of every record, memo_cli.py:131-135,:158-167): the text bytes are uploaded in chunks
 * of whole records, one warp per record tokenises ([a-z0-9_]+ after ASCII lower-casing; lower-case non-ASCII text
 * with Unicode rules first), hashes every token (CPython SipHash-1-3, PYTHONHASHSEED=0 key), accumulates the
 * signed buckets in shared memory, normalises with K1's arithmetic (normalize != 0) and stores the row straight
 * into the index.  skip_blank != 0 leaves out records that hold nothing but ASCII white space, as the reference's
 * rebuild does; kept rows stay in record order.  ids: ids_host[i], or first_id + i (the record's position) when
 * ids_host is NULL and with_ids != 0, or none (add() semantics) when both are 0.  n_added receives the rows added. */
int b200_index_add_texts(b200_index* ix, const char* utf8_host, const int64_t* offsets_host, int64_t n,
                         const int64_t* ids_host, int64_t first_id, int skip_blank, int normalize, int with_ids,
                         int64_t* n_added);
/* counter-based synthetic rows written to a device buffer [n,d] float32 */
int b200_synth_rows_dev(float* out_dev, int64_t n, int d, uint64_t seed, int64_t first_row,
                        int normalize, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_FLAT_H */
