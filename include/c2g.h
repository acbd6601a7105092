/*
 * c2g.h — the C-ABI of the B200-native cont2contops hot path (libc2g.so).
 *
 * The reference has no FFI/plugin interface: its "plugin API" is the C++ class surface ContourManager / ContourDB that
 * test/batch_bin_test.cpp drives.  This header is the thin boundary a maintainer binds to keep that
 * surface and move the work to the GPU: plain pointers, sizes and POD structs (include/c2g_types.h), no torch / STL
 * types, `int` return codes (0 = ok, < 0 = -cudaError_t, <= -1000 = argument error), never throws.
 * The host-side C++ facade in contour_context_b200/host/ (same class and method names as the reference) and the
 * Python mirror in contour_context_b200/ are both written against exactly these entry points.
 *
 * Each entry point cites the reference interface it replaces (file:line relative to the reference repo).
 * One context per (process, device); calls are stream-ordered on the context's stream; not thread-safe.
 */
#ifndef C2G_H
#define C2G_H

#include "c2g_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct c2g_ctx c2g_ctx;

#define C2G_ERR_ARG (-1000)
#define C2G_ERR_CAPACITY (-1001)
#define C2G_ERR_STATE (-1002)

/* Library / ABI version and struct sizes (lets a binding verify it was generated against the same header). */
int c2g_abi_version(void);
int c2g_sizeof(int which); /* 0 scan_head, 1 view, 2 bci, 3 hint, 4 pair_score, 5 query_result, 6 cm_config, 7 db_config */

/* Replaces: ContourManager::ContourManager(cfg, int_id) storage (include/cont2/contour_mng.h:478-498) and
 * ContourDB::ContourDB(cfg) (include/cont2/contour_db.h:680-684).
 * `scan_capacity` = number of scan slots (DB scans + in-flight query scans) resident in HBM,
 * `max_batch` = largest number of scans one c2g_ingest / c2g_query call may carry,
 * `max_points` = largest total number of points of one batch (staging buffer for host inputs). */
int c2g_create(const c2g_cm_config *cm_cfg, const c2g_db_config *db_cfg, int device, int scan_capacity, int max_batch,
               long long max_points, c2g_ctx **out);
int c2g_destroy(c2g_ctx *ctx);

/* Use an externally owned CUDA stream (cudaStream_t passed as void*) for all subsequent work; NULL = context's own. */
int c2g_set_stream(c2g_ctx *ctx, void *cuda_stream);
int c2g_sync(c2g_ctx *ctx);

/* Replaces, for a batch of B scans: readKITTIPointCloudBin's output buffer (include/tools/pointcloud_util.h:12-50,
 * N x 4 float32 per scan) -> ContourManager::makeBEV (include/cont2/contour_mng.h:505-556) ->
 * ContourManager::makeContoursRecurs (:588-960).  Scan b's points are pts[4*offsets[b] .. 4*offsets[b+1]).
 * `pts_on_device` != 0: pts is a device pointer (16-byte aligned); otherwise a host pointer that is copied through the
 * context's staging buffer inside the call (pinned host memory makes that copy asynchronous).
 * The descriptors land in scan slots first_slot .. first_slot + B - 1; int_ids (host, may be NULL) are the
 * ContourManager int ids (evaluator.h:290). Asynchronous on the context stream. */
int c2g_ingest(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device, int first_slot,
               const int *int_ids_host);

/* The same with 12 bytes per point: scan b's points are xyz[3*offsets[b] .. 3*offsets[b+1]) as (x, y, z).  The reference reads the
 * fourth float of the KITTI .bin record (intensity) and never uses it (include/tools/pointcloud_util.h:17-30, makeBEV reads
 * x, y, z only: contour_mng.h:505-556); a caller that drops it on the host moves 25 % fewer bytes over PCIe.  Results are
 * identical to c2g_ingest on the same points.  (Separately reported variant of SURVEY.md 8f-4; device pointers need 4-byte
 * alignment only.) */
int c2g_ingest_xyz(c2g_ctx *ctx, const float *xyz, const long long *offsets_host, int B, int pts_on_device, int first_slot,
                   const int *int_ids_host);

/* Page-locked host memory for point buffers handed to c2g_ingest (the buffer readKITTIPointCloudBin fills,
 * include/tools/pointcloud_util.h:17-30): the host -> device copy of a pinned buffer runs at PCIe speed and asynchronously,
 * a pageable one is bounced through driver staging. Plumbing only (cudaHostAlloc / cudaFreeHost). */
int c2g_host_alloc(void **out, size_t bytes);
int c2g_host_free(void *p);

/* First stage of c2g_ingest alone (the BEV scatter kernel: makeBEV, contour_mng.h:505-556), for profiling and parity tests. */
int c2g_ingest_bev_only(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device);
/* What the scatter kernel hands to the contour kernel for scan `batch_index` of the last batch (parity tests): the six
 * bit-planes `bev > lv_grads[l]` (cv::threshold of makeContourRecursiveHelper, src/cont2/contour_mng.cpp:283; C2G_NLEV x n_row x
 * ceil(n_col / 32) words, bit c & 31 of word r * ceil(n_col / 32) + (c >> 5)), the cells above the lowest threshold in raster
 * order as (height, row_f, col_f, 0) (bev_ + bev_pixfs_, contour_mng.h:435,528-529; fg_host must hold 4 * n_row * n_col floats)
 * and hdr_host[2] = (occupied cells = bev_pixfs_.size(), foreground cells).  full_tile_variant != 0: recomputed by the full-tile
 * variant of the kernel (the one behind c2g_get_bev), which must agree bit for bit with the production variant. */
int c2g_get_bev_compact(c2g_ctx *ctx, int batch_index, int full_tile_variant, unsigned int *planes_host, float *fg_host, int *hdr_host);

/* Read back (synchronises the stream): ContourManager getters (include/cont2/contour_mng.h:1052-1106). */
int c2g_get_heads(c2g_ctx *ctx, int first_slot, int n, c2g_scan_head *out_host);
int c2g_get_views(c2g_ctx *ctx, int slot, c2g_view *out_host /* C2G_VIEW_CAP records */);
/* Dense BEV of scan `batch_index` of the LAST ingest batch: ContourManager::getBevImage (:573-586) plus the
 * continuous pillar coordinates bev_pixfs_ (:435); empty cells are -1000 / -1 / -1. Each array holds n_row*n_col.
 * The image is produced on demand by re-running the scatter kernel's full-tile variant on that one scan (the ingest
 * kernels only materialise the cells above the lowest threshold), so a device-resident points buffer handed to the last
 * c2g_ingest must still be alive. */
int c2g_get_bev(c2g_ctx *ctx, int batch_index, float *bev, float *row_f, float *col_f);
/* Raw 64-bit BEV cell keys of the last batch (debug / K1 parity). */
int c2g_get_tiles(c2g_ctx *ctx, int batch_index, unsigned long long *out_host);

/* Copy finished descriptors between slots (ContourDB::addScan keeps the shared_ptr, include/cont2/contour_db.h:823). */
int c2g_copy_slots(c2g_ctx *ctx, int src_first, int dst_first, int n);

/* Replaces ContourDB::addScan (include/cont2/contour_db.h:814-824) for the n finished scans in slots first_slot.. (which
 * must equal the current DB size: gidx == slot == all_bevs_.size()): their non-zero q-level keys enter the time-delay
 * buffers of the host-side LayerDBs with timestamps ts_host[i]. Synchronises the stream (reads the keys back, 1.4 KB/scan). */
int c2g_db_add_scans(c2g_ctx *ctx, int first_slot, int n, const double *ts_host);
/* Replaces ContourDB::pushAndBalance (include/cont2/contour_db.h:827-843, LayerDB::rebuild src/cont2/contour_db.cpp:63-317). */
int c2g_db_push_and_balance(c2g_ctx *ctx, int seed, double ts);
int c2g_db_size(c2g_ctx *ctx);

/* Windowed online loop: replaces W consecutive iterations of BatchBinSpinner::spinOnce's database half
 * (test/batch_bin_test.cpp:179,234,237), for i = 0 .. W-1 in order:
 *     ContourDB::queryRangedKNN(scan_i)   include/cont2/contour_db.h:698-811
 *     ContourDB::addScan(scan_i, ts[i])   include/cont2/contour_db.h:814-824
 *     ContourDB::pushAndBalance(seeds[i], ts[i])   include/cont2/contour_db.h:827-843
 * with results identical to the scan-by-scan calls (query i sees exactly the trees that exist after scans < i were added
 * and balanced): the W scans are ingested in one batch into slots db_size .. db_size + W - 1, the host replays the LayerDB
 * bookkeeping of the window from their keys and records for every scan the state of the trees it must see (bucket boundaries,
 * sizes, regions); the device mirror keeps those states readable (appends + a second region per bucket for rewrites), so the
 * window's patches are applied first and ONE kNN launch serves all its scans; the rest of the query chain runs once for the
 * window.
 *   c2g_online_stage   ingests the next window (arguments as c2g_ingest) and starts the read-back of its keys; asynchronous.
 *                      At most two windows may be staged: stage window k+1, then commit window k, and the host -> device copy
 *                      of k+1 overlaps the bookkeeping and the query kernels of k.
 *   c2g_online_commit  oldest staged window: bookkeeping + queries; results_host[W] (pinned memory recommended) is filled
 *                      asynchronously: valid after c2g_sync.
 *   c2g_online_window  stage + commit + sync in one call.
 *   c2g_online_runs    runs of scans that saw identical trees in the windowed loop so far;
 *   c2g_online_groups  kNN launches it issued (one launch serves a group of runs: every scan carries its own view of the trees). */
int c2g_online_stage(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host);
/* the same for 12 B / point buffers (x, y, z), as c2g_ingest_xyz */
int c2g_online_stage_xyz(c2g_ctx *ctx, const float *xyz, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host);
int c2g_online_commit(c2g_ctx *ctx, const double *ts_host, const int *seeds_host, const c2g_score_ensemble *lb,
                      const c2g_score_ensemble *ub, c2g_query_result *results_host);
int c2g_online_window(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host,
                      const double *ts_host, const int *seeds_host, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
                      c2g_query_result *results_host);
long long c2g_online_runs(c2g_ctx *ctx);
long long c2g_online_groups(c2g_ctx *ctx);
/* Measurement aid: host seconds c2g_online_commit has spent so far in [0] the LayerDB bookkeeping, [1] kNN launches, [2] mirror
 * patches, [3] the launches of the rest of the chain. */
int c2g_online_host_seconds(c2g_ctx *ctx, double *out4);
/* Upload the host-side tree contents to the device tables if they changed (c2g_query* call it implicitly). */
int c2g_db_sync(c2g_ctx *ctx);
/* Introspection for parity tests: bucket boundaries [7], tree sizes [6], buffer sizes [6]; one bucket's tree in order. */
int c2g_db_layer_state(c2g_ctx *ctx, int ll, float *bucket_ranges, int *tree_sizes, int *buffer_sizes);
int c2g_db_bucket_tree(c2g_ctx *ctx, int ll, int bucket, float *keys, int *gidx, int *seq);

/* Low-level mirror call used by c2g_db_sync: device copy of the host-maintained LayerDB state (include/cont2/contour_db.h:159-217, src/cont2/contour_db.cpp:63-317):
 * for q-level index `ll`, the keys currently INSIDE the KD-trees (not the time-delay buffers), the bucket each one
 * lives in, where it came from (IndexOfKey gidx/seq) and the 7 bucket boundaries. keys: n x 10 floats, row-major. */
int c2g_db_set_layer(c2g_ctx *ctx, int ll, int n, const float *keys_host, const int *gidx_host, const signed char *seq_host,
                     const unsigned char *bucket_host, const float *bucket_ranges_host);

/* Replaces ContourDB::queryRangedKNN (include/cont2/contour_db.h:698-811) for B query scans sitting in slots
 * first_slot.. : ranged kNN over the mirrored key tables (LayerDB::layerKNNSearch, contour_db.cpp:319-379) ->
 * CandidateManager::checkCandWithHint for every hint (contour_db.h:374-488) -> tidyUpCandidates (:494-596) ->
 * fineOptimize's selection without the Ceres step (:604-648). gidx of a DB scan == its slot.
 * results_host: B records. hints_host / scores_host (optional, may be NULL): B * n_q_levels * piv * nnk records in
 * reference order (layer, query seq, ascending distance), empty slots have cand_gidx = -1. Synchronises the stream. */
int c2g_query(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
              c2g_query_result *results_host, c2g_hint *hints_host, c2g_pair_score *scores_host);
/* Same, asynchronous, results stay on the device (for multi-GPU gathers and timing); pointers are device pointers
 * owned by the context, valid until the next query call. */
int c2g_query_async(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub);
int c2g_query_buffers(c2g_ctx *ctx, void **results_dev, void **hints_dev, void **scores_dev, long long *n_hint_slots);
/* Asynchronous export of the last query's buffers: hint / pair-score records to caller-owned DEVICE buffers (the send
 * buffers of the multi-GPU all-gather) and/or the per-query results to a (pinned) HOST buffer. NULL skips a part. */
int c2g_query_export(c2g_ctx *ctx, int B, void *hints_dst_dev, void *scores_dst_dev, void *results_dst_host);
/* Finish a query from pair-score records produced elsewhere (multi-GPU: after the all-gather of score records):
 * replay CandidatePoseData::addProposal / tidyUpCandidates for B queries from device arrays laid out like
 * c2g_query_buffers'. */
int c2g_finish_from_scores(c2g_ctx *ctx, int first_slot, int B, const c2g_score_ensemble *lb, const void *hints_dev,
                           const void *scores_dev, c2g_query_result *results_host);

/* Stand-alone handles on the host-side LayerDB logic (no CUDA context): used by the C++ facade's unit tests and the
 * CPU parity tests of the bucket rebalancing (src/cont2/contour_db.cpp:63-317). */
void *c2g_hostdb_create(int n_layers, double max_elapse, double min_elapse);
void c2g_hostdb_free(void *h);
int c2g_hostdb_push_key(void *h, int ll, const float *key, double ts, int gidx, int seq);
int c2g_hostdb_balance(void *h, int seed, double ts);
int c2g_hostdb_state(void *h, int ll, float *bucket_ranges, int *tree_sizes, int *buffer_sizes);
int c2g_hostdb_tree(void *h, int ll, int bucket, float *keys, int *gidx, int *seq);
/* searchable prefix of every bucket's tree (what the reference's KD index covers: rebuilt only when the bucket pops its buffer) */
int c2g_hostdb_indexed(void *h, int ll, int *indexed /* [6] */);
/* per bucket: how often existing tree entries were moved or removed (LayerDB::rebuild's balancing move); while the counter
 * stands still a tree only grows at its END, which is what lets c2g_db_sync patch the device mirror instead of rebuilding it */
int c2g_hostdb_versions(void *h, int ll, unsigned int *restructured /* [6] */);

/* Which exp() variant the device runs to match this host's libm (csrc/c2g_libm.cuh): 0 libdevice, 1 glibc, 2 glibc+FMA. */
int c2g_exp_mode(c2g_ctx *ctx);
/* Host execution of the libm restatements (tests only). */
int c2g_selftest_libm(int kind, int n, const void *in, void *out);

/* Introspection for tests: how many scans of the LAST scatter launch (the last chunk of the last c2g_ingest* / c2g_bev_only call) the
 * fast scatter kernel handed to the general 64-bit kernel (more than 2^17 points, event log or foreground list overflow). */
int c2g_scatter_deferred(c2g_ctx *ctx, int *n_out);

/* Developer aid: clock64() stamps of the contour kernel's phases (64 values). */
int c2g_debug_clocks(c2g_ctx *ctx, long long *out_host);

/* Counters: kernels launched by this context since creation (bench.py's gpu_launches). */
long long c2g_launch_count(c2g_ctx *ctx);
/* Measurement aid (no reference counterpart): with enable != 0, c2g_query_async brackets each of its kernels with CUDA events
 * on the context's stream.  When ms_out != NULL the durations (ms) of the last profiled query are written first:
 * ms_out[8] = knn, prefilter, score, proposal replay, GMM-L2 gate, output, refinement, ranking.  Synchronises the stream. */
int c2g_query_profile(c2g_ctx *ctx, int enable, float *ms_out);

/* Measurement aid (no reference counterpart): work actually done by the query kernels since the last call, for the roofline
 * entries of bench.py.  out_host[8] (may be NULL) = kNN keys distance-evaluated, kNN block boxes tested, GMM-L2 gate pre-selection
 * tests (correlation.h:85-96), GMM-L2 gate Gaussian terms (:138-149), refinement pre-selection tests, refinement Gaussian terms
 * (pairs x cost+gradient evaluations), refinement evaluations, spare.  Reads and clears the counters, then enables / disables
 * counting (one atomic per warp while enabled).  Synchronises the stream. */
int c2g_work_counters(c2g_ctx *ctx, int enable, unsigned long long *out_host);

/* Host-side replay of libstdc++ std::sort used by the kernels (tests only): sorts `n` packed (key << 16 | index)
 * words with comparator key-descending (desc != 0) or key-ascending. */
int c2g_selftest_stdsort(unsigned int *words, int n, int desc);
/* The warp-cooperative replay of the same std::sort that the contour kernel runs (csrc/stdsort.cuh), on the device (tests only). */
int c2g_selftest_warpsort(c2g_ctx *ctx, unsigned int *words, int n, int desc);

#ifdef __cplusplus
}
#endif
#endif /* C2G_H */
