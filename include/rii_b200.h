/* rii_b200.h -- C ABI of librii_b200.so: the B200 (sm_100a) implementation of Rii's ADC hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one member of the reference's pybind11 class
 * `main.RiiCpp` (matsui528/rii v0.2.12, src/main.cpp:12-54; implementation src/rii.h) and is what a
 * maintainer's FFI stub (ctypes / pybind11 / cgo) binds -- see INTEGRATION.md.  Plain pointers and sizes
 * only; no C++/torch types cross the boundary.  All functions return 0 (or a non-negative count) on
 * success and a negative code on failure, with the message available from rii_last_error(); nothing
 * throws and nothing aborts (the reference asserts/terminates, src/rii.h:166-170).
 *
 * Pointers named `h_*`/unprefixed are HOST pointers; `d_*` are DEVICE pointers on the index's GPU.
 * Results use the total order (distance ascending, id ascending); ids are global 64-bit ids.
 */
#ifndef RII_B200_H
#define RII_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rii_index rii_index_t;

#define RII_OK 0
#define RII_ERR_ARG (-1)      /* invalid argument (the reference would assert) */
#define RII_ERR_STATE (-2)    /* call not valid in this state (e.g. update without coarse centers) */
#define RII_ERR_CUDA (-3)     /* CUDA runtime error */
#define RII_ERR_LIMIT (-4)    /* shape outside the implemented range */

#define RII_METHOD_LINEAR 0
#define RII_METHOD_IVF 1

/* Message of the last failure on the calling thread ("" if none). */
const char *rii_last_error(void);
/* "0.2.12+b200.<n>": tracks main.__version__ (src/main.cpp:56-60). */
const char *rii_version(void);
/* Number of kernels this library launched on behalf of the calling process (bench.py gpu_launches). */
int64_t rii_launch_count(void);

/* RiiCpp(codewords, verbose)  src/main.cpp:14, src/rii.h:86-106.
 * codewords: float32 (M, Ks, Ds) row-major, copied.  Ks <= 256 (rii/rii.py:35).  device: CUDA ordinal.
 * l2_variant: accumulator width of the reference build whose fvec_L2sqr rounding is reproduced
 * (src/distance.h:113,172,219): 16 = AVX-512, 8 = AVX, 4 = SSE, 0 = pick from this host's CPU flags.
 * (The variants differ only when Ds >= 16.) */
int rii_create(const float *codewords, int M, int Ks, int Ds, int verbose, int device, int l2_variant,
               rii_index_t **out);
int rii_destroy(rii_index_t *h);

/* RiiCpp::add_codes(codes, update_flag)  src/main.cpp:16, src/rii.h:158-193.  codes: uint8 (n, M). */
int rii_add_codes(rii_index_t *h, const uint8_t *codes, int64_t n, int update_flag);
/* Same with the codes already in DEVICE memory on the index's GPU (index builds at 1e8-1e9 codes: no host round trip).
 * Synchronises the device first (the producer of d_codes may run on any stream), like rii_set_lists_dev. */
int rii_add_codes_dev(rii_index_t *h, const uint8_t *d_codes, int64_t n, int update_flag);
/* RiiCpp::reconfigure(nlist, iter)  src/main.cpp:15, src/rii.h:108-156 (+ PQk-means src/pqkmeans.cpp). */
int rii_reconfigure(rii_index_t *h, int nlist, int iter);
/* RiiCpp::clear()  src/main.cpp:28, src/rii.h:328-333. */
int rii_clear(rii_index_t *h);

/* RiiCpp::query_linear(query, topk, target_ids)  src/main.cpp:17-21, src/rii.h:195-242.
 * query: float32 (M*Ds); target_ids: int64 (S) or NULL/S=0 for all.  Writes <= topk results, returns the
 * count (== topk) or a negative error. */
int64_t rii_query_linear(rii_index_t *h, const float *query, int topk, const int64_t *target_ids, int64_t S,
                         int64_t *out_ids, float *out_dists);
/* RiiCpp::query_ivf(query, topk, target_ids, L)  src/main.cpp:22-27, src/rii.h:244-326.
 * target_ids must be sorted ascending (src/rii.h:294).  Returns 0 for the reference's empty result
 * (src/rii.h:325). */
int64_t rii_query_ivf(rii_index_t *h, const float *query, int topk, const int64_t *target_ids, int64_t S, int64_t L,
                      int64_t *out_ids, float *out_dists);
/* Batch entry (new: the reference takes one query per call, rii/rii.py:251).  queries: float32 (B, M*Ds);
 * one shared target_ids set; method RII_METHOD_*; L ignored for linear.  out_ids/out_dists: (B, topk),
 * out_counts: (B) results per query. */
int rii_query_batch(rii_index_t *h, const float *queries, int B, int topk, const int64_t *target_ids, int64_t S,
                    int64_t L, int method, int64_t *out_ids, float *out_dists, int32_t *out_counts);
/* Same with DEVICE buffers, enqueued on `stream` (a cudaStream_t; NULL = the CUDA default stream) and not
 * synchronised unless the rare full-ranking re-run (walk beyond w) is needed.  The handle's scratch
 * buffers are shared: one call in flight per handle (calls on different streams must be ordered by the caller; the
 * derived code layouts a call builds lazily are guarded by an event, so a later call on another stream waits for them).
 * d_target_ids: device int64 (S) or NULL. */
int rii_query_batch_dev(rii_index_t *h, const float *d_queries, int B, int topk, const int64_t *d_target_ids,
                        int64_t S, int64_t L, int method, int64_t *d_out_ids, float *d_out_dists,
                        int32_t *d_out_counts, void *stream);

/* Properties  src/main.cpp:29-34. */
int64_t rii_get_N(const rii_index_t *h);
int rii_get_nlist(const rii_index_t *h);
int rii_get_verbose(const rii_index_t *h);
int rii_set_verbose(rii_index_t *h, int verbose);
int rii_get_dims(const rii_index_t *h, int *M, int *Ks, int *Ds);
/* flattened_codes -> (N, M) uint8; coarse_centers -> (nlist, M) uint8; posting_lists -> CSR
 * (offsets int64 (nlist+1), ids int32 (N)).  Caller allocates. */
int rii_copy_codes(const rii_index_t *h, uint8_t *out);
int rii_copy_coarse_centers(const rii_index_t *h, uint8_t *out);
int rii_copy_posting_lists(const rii_index_t *h, int64_t *offsets, int32_t *ids);
/* __setstate__  src/main.cpp:39-53: replace codes / coarse centers / posting lists wholesale. */
int rii_set_state(rii_index_t *h, const uint8_t *coarse_centers, int nlist, const uint8_t *codes, int64_t N,
                  const int64_t *offsets, const int32_t *ids);

/* ---- building blocks of the path, exposed for parity tests and multi-GPU orchestration ---------- */
/* K1: distance tables of B queries -> out (B, M, Ks) float32 (host).  src/rii.h:361-373. */
int rii_dtable(rii_index_t *h, const float *queries, int B, float *out);
/* ADC distance of every stored code to one query -> out (N) float32 (host).  src/rii.h:386-394. */
int rii_adist_all(rii_index_t *h, const float *query, float *out);
/* K6: nearest coarse center (symmetric distance, first minimum wins) of n codes against K centers.
 * src/pqkmeans.cpp:193-218.  out_assign int32 (n), out_dist float32 (n) or NULL. */
int rii_assign(rii_index_t *h, const uint8_t *codes, int64_t n, const uint8_t *centers, int K, int32_t *out_assign,
               float *out_dist);
/* Codeword distance matrices (M, Ks, Ks) float32.  src/pqkmeans.cpp:23-34. */
int rii_sym_matrices(rii_index_t *h, float *out);

/* PQ encoder (the step before the path: fine_quantizer.encode, rii/rii.py:185): vecs float32 (n, M*Ds) ->
 * out_codes uint8 (n, M); nearest codeword per sub-space, fp32 sequential sum, first minimum wins. */
int rii_encode(rii_index_t *h, const float *vecs, int64_t n, uint8_t *out_codes);

/* OPQ: rotate every query on the device before the distance table is built (rii/rii.py:305-306 does fine_quantizer.rotate
 * on the host).  R: float32 (D, D) row-major, q' = q @ R; NULL switches it off.  fp32 FMA chain: agrees with a numpy
 * rotation to rounding (~1e-7 relative), not bit for bit. */
int rii_set_rotation(rii_index_t *h, const float *R);

/* ---- id-range sharding (one index object per GPU) -------------------------- */
/* This shard holds global ids [id_base, id_base + N_local) of an index of N_total codes. */
int rii_set_shard(rii_index_t *h, int64_t id_base, int64_t N_total);
/* Replace coarse centers and assign every local code (UpdatePostingLists(0, N), src/rii.h:335-359). */
int rii_set_coarse_centers(rii_index_t *h, const uint8_t *centers, int nlist);
/* PQk-means on a given sample (src/pqkmeans.cpp:46-133): sample uint8 (ns, M) already in the reference's
 * shuffled order -> centers_out (nlist, M). */
int rii_fit_coarse(rii_index_t *h, const uint8_t *sample, int64_t ns, int nlist, int iter, uint8_t *centers_out);
/* Local list lengths (nlist) int32. */
int rii_copy_list_lengths(const rii_index_t *h, int32_t *out);
/* Global list lengths and the lengths held by lower ranks (both (nlist) int32), from the all-gather. */
int rii_set_global_lengths(rii_index_t *h, const int32_t *glob_len, const int32_t *pre_len);

/* Posting lists from an EXTERNAL clustering: coarse centers (host, (nlist, M)) and the list of every local row (device
 * int32 (N), values in [0, nlist)); the lists are built on the device (ascending ids per list), K6 is skipped. */
int rii_set_lists_dev(rii_index_t *h, const uint8_t *centers, int nlist, const int32_t *d_assign);
/* Pre-allocate the code table for `rows` codes (large builds: no regrowth copies). */
int rii_reserve(rii_index_t *h, int64_t rows);
/* First min(N_total, 100*nlist) ids of the reference's sampling shuffle (src/rii.h:115-124; host only).
 * Call with out_ids == NULL to get the count in *out_n. */
int rii_sample_ids(int64_t N_total, int nlist, int64_t *out_ids, int64_t *out_n);
/* Merge G per-shard results (device buffers: ids int64 (G, B, k) global ids, dists float32 (G, B, k), counts
 * int32 (G, B)) into the global top-k per query under (distance, id).  The buffers are what an all-gather of
 * rii_query_batch_dev outputs produces. */
int rii_merge_shards_dev(rii_index_t *h, const int64_t *d_ids, const float *d_dists, const int32_t *d_counts, int G,
                         int B, int k, int64_t *d_out_ids, float *d_out_dists, int32_t *d_out_counts, void *stream);

/* The same merge over ONE packed buffer (a single all-gather): shard g's block starts at g * stride_bytes and holds
 * [ids int64 (B, k) | dists float32 (B, k) | counts int32 (B)]; stride_bytes must be a multiple of 8. */
int rii_merge_shards_packed_dev(rii_index_t *h, const void *d_packed, int64_t stride_bytes, int G, int B, int k, int64_t *d_out_ids,
                                float *d_out_dists, int32_t *d_out_counts, void *stream);

/* IVF + target_ids on an id-range shard (the binary_search filter of src/rii.h:294 with the ids spread over shards):
 * the cut after L member candidates and the topk test at the w-th list are global, so every shard needs the member
 * counts of the others -- per list.  Per target set, on every shard:
 *   rii_subset_begin_dev(d_target_ids (sorted ascending, global ids), S, d_list_counts (nlist) out)
 *        builds the sub-index of this shard's members and returns its per-list member counts
 *   [caller: all-gather of the counts; d_counts_all = sum over shards, d_counts_lower = sum over lower ranks]
 *   rii_subset_set_global_dev(d_counts_all, d_counts_lower)        (skip on a single shard)
 *   rii_subset_query_dev(queries ...)   any number of times: an ordinary sharded IVF search over the sub-indexes, S taken
 *        from begin; per-shard top-k out (then all-gather + rii_merge_shards_dev).  Queries that walk beyond the first w
 *        lists (src/rii.h:309-322) are re-run with the full ranking inside the call; count 0 = the reference's empty
 *        result (src/rii.h:325).
 * The prepared sub-index stays valid until the next begin / subset query through rii_query_batch* / index update. */
int rii_subset_begin_dev(rii_index_t *h, const int64_t *d_target_ids, int64_t S, int32_t *d_list_counts, void *stream);
int rii_subset_set_global_dev(rii_index_t *h, const int32_t *d_counts_all, const int32_t *d_counts_lower, void *stream);
int rii_subset_query_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, int64_t *d_out_ids, float *d_out_dists,
                         int32_t *d_out_counts, void *stream);

/* The coarse phase of IVF (src/rii.h:259-280) split from the posting-list scan, for sharded batches: every rank ranks the
 * lists for ITS share of the batch (the ranking does not depend on the shard), the (B, w) rankings are all-gathered, and
 * every rank scans its shard for all queries with the rankings given.  w = rii_coarse_width(h, L).
 * d_flags (B, may be NULL): bit 0 = fewer than topk candidates in the first w lists (re-run through rii_query_batch_dev,
 * which ranks all lists), bit 1 = empty result.  Streaming-engine shapes only (12 <= M <= 64, topk <= 224). */
int rii_coarse_width(rii_index_t *h, int64_t L);
int rii_coarse_rank_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, int32_t *d_ranked, void *stream);
int rii_query_ranked_dev(rii_index_t *h, const float *d_queries, int B, int topk, int64_t L, const int32_t *d_ranked, int64_t *d_out_ids,
                         float *d_out_dists, int32_t *d_out_counts, int32_t *d_flags, void *stream);

/* Tuning knobs.  "scan_kernel": 0 = auto, 1 = natural-layout kernels only, 4 = the skew64 streaming engine or fail
 * (what auto picks for 12 <= M <= 64, topk <= 224, ascending target_ids, N >= 32768).  "persist": 0 = never use the persistent
 * warp-specialised batch kernel, 1 = auto (batches of >= 296 queries), 2 = whenever the shape fits.  "stream_ctas": 1 = always
 * one CTA per SM.  "fuse_coarse": 0 = coarse ranking in its own launch.  "zero_copy": 0 = small host calls use staged copies
 * instead of the pinned, device-mapped buffer.  "assign_kernel": 0 = auto, 1 = natural-layout k_assign, 3 = streaming
 * assignment engine with one CTA per SM.  "debug_clocks": 1 = record per-CTA phase clocks (rii_debug_clocks).
 * Results are identical for every setting. */
int rii_set_option(rii_index_t *h, const char *name, int64_t value);

/* ---- measurement ------------------------------------------------------------------------------ */
/* Per-kernel device time from CUDA events recorded around every launch on the launching stream.
 * kernel: "dtable" | "scan_linear" | "merge" | "coarse_rank" | "subset_build" | "plan" | "scan_ivf" | "assign" | "sort". */
int rii_profile_enable(rii_index_t *h, int on);
int rii_profile_reset(rii_index_t *h);
int rii_profile_get(rii_index_t *h, const char *kernel, double *ms_total, int64_t *launches);
/* With option "debug_clocks" = 1 the scan kernels record clock64() values per CTA; this copies (n_ctas, 8) int64 of the last
 * launch to the host.  k_scan_stream32: [0] start, [1] ready to scan (table, and for the fused kernel coarse ranking + plan),
 * [2] scan done, [3] end, [4] table built, [5] coarse pass done, [6] coarse pool sorted, [7] = pool size.  k_scan_persist32:
 * 16 counters per CTA (two rows): cycles waiting / scanning of consumer warps 0 and 10, producer cycles waiting / merging /
 * table / coarse pass / selection / plan, queries of the CTA (tools/phase_clocks.py). */
int rii_debug_clocks(rii_index_t *h, int64_t n_ctas, int64_t *out);

#ifdef __cplusplus
}
#endif
#endif /* RII_B200_H */
