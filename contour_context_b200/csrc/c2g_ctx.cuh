// c2g_ctx.cuh — the context object behind the opaque c2g_ctx handle of include/c2g.h.
#pragma once
#include "c2g_common.cuh"
#include <time.h>
#define C2G_TRACE_CAP 4096

#define C2G_MAX_CHUNK_EVENTS 64
#define C2G_QUERY_STREAMS 8  // sub-batches of one c2g_query_async call that may run concurrently
#define C2G_WORK_N 8   // knn keys evaluated, knn block boxes tested, gate pre-selection tests, gate terms, refine pre-selection tests,
                       // refine pair terms (pairs x evaluations), refine evaluations, (spare)
#define C2G_PATCH_RING 2048
#define C2G_QPROF_N 9  // knn, prefilter, score, replay, corr, output, refine, rank, (spare)

#define C2G_PHYS_BUCKETS (2 * C2G_NUM_BUCKETS)
struct C2gLayerTable {   // device mirror of one LayerDB's KD-tree contents (host logic keeps the authoritative copy)
  // Bucket k owns two fixed regions p = k, k + C2G_NUM_BUCKETS: [p * cap_b, (p + 1) * cap_b) of every per-key array and [p * blkcap_b, ..)
  // of the block arrays.  phys[k] is the one in use: a change in one bucket never moves the entries of another, appending to
  // a bucket never disturbs a kNN launch that reads its shorter prefix, and a REWRITE of a bucket (rebalancing move, kd
  // re-ordering) goes to the bucket's other region, so that launches still reading the previous version are not disturbed
  // either (the windowed online loop runs the kNN of a whole window after all of the window's patches, c2g_online_commit).
  float *keys_t;         // [C2G_KEY_DIM][C2G_PHYS_BUCKETS * cap_b] transposed for coalesced scans
  int *gidx;             // IndexOfKey::gidx
  signed char *seq;      // IndexOfKey::seq
  int *orank;            // region base + position in TREE order (tie-break rank; the mirror itself may be kd-ordered)
  float *box_min, *box_max;  // [C2G_KEY_DIM][C2G_PHYS_BUCKETS * blkcap_b]: bounding box of every 32-key block
  int cap_b, blkcap_b;
  int phys[C2G_NUM_BUCKETS];
  int bucket_cnt[C2G_NUM_BUCKETS];
  float ranges[C2G_NUM_BUCKETS + 1];
  // what the mirror currently holds per bucket (c2g_db_sync patches the difference to the host trees)
  int m_n[C2G_NUM_BUCKETS];             // entries mirrored
  unsigned m_rv[C2G_NUM_BUCKETS];       // C2gBucket::restructured at that time
  unsigned char m_kd[C2G_NUM_BUCKETS];  // mirrored in kd-blocked order (batch queries) or in tree order (online loop)
  unsigned char m_valid[C2G_NUM_BUCKETS];
};

struct C2gHostDB;  // host-side LayerDB state (layer_db_host.h)

struct c2g_ctx {
  int device, num_sms;
  C2gIngestParams P;
  c2g_db_config db;
  int scan_cap, max_batch;
  long long max_points;
  cudaStream_t own_stream, stream;
  // ingest buffers
  float *d_pts_stage2[2];              // double-buffered staging of host point batches
  int stage_sel;
  cudaStream_t copy_stream;            // H2D of chunk k+1 overlaps the kernels of chunk k
  cudaEvent_t ev_stage_free[2];        // recorded when the kernels reading a staging buffer are done
  cudaEvent_t ev_chunk[C2G_MAX_CHUNK_EVENTS];
  long long *d_offsets;                // offsets / internal ids of a batch whose points came by device pointer or c2g_bev_only
  int *d_int_ids;
  long long *d_offsets2[2];            // the same per staging buffer of the pipelined host ingest (uploaded on the copy stream)
  int *d_int_ids2[2];
  const long long *last_offsets;       // offsets of the last batch (c2g_get_bev re-reads them)
  cudaEvent_t ev_patch_up;             // a window's mirror patches have arrived (copy stream)
  // scatter kernel -> contour kernel hand-off of one batch: bit-planes, foreground cell lists, (occupied, foreground) counts
  uint32_t *d_planes;      // [max_batch][C2G_NLEV][n_row * ceil(n_col / 32)]
  float4 *d_fg;            // [max_batch][n_cells]
  int2 *d_hdr;             // [max_batch]
  int *d_work_counter_k1;  // scatter kernel launch: [0] next scan, [1] next deferred scan, [2] deferred count, [3..] deferred scans
  // one-scan scratch of the dense-image getters (c2g_get_bev / c2g_get_tiles: the full-tile scatter variant run on demand)
  c2g_cellkey *d_tile1;
  uint32_t *d_planes1;
  float4 *d_fg1;
  int2 *d_hdr1;
  float *d_bev_h, *d_bev_rf, *d_bev_cf;  // [n_cells] each
  c2g_view *d_presort;
  c2g_scan_head *d_heads;
  c2g_view *d_views;
  c2g_ell *d_ells;           // [scan_cap][C2G_VIEW_CAP], same indexing as d_views
  long long *d_dbg;
  int *d_work_counter;     // next scan of the running contour_kernel launch
  unsigned char *d_k2_scratch;  // contour kernel: per resident CTA key-window lists + overflow arenas (contours.cu)
  const float *last_pts;
  int last_B;
  int last_fpp;  // floats per point of the last batch: 4 (KITTI .bin layout) or 3 (c2g_ingest_xyz)
  long long launches;
  // query buffers
  C2gLayerTable layers[C2G_NUM_Q_LEVELS_MAX];
  c2g_hint *d_hints;
  c2g_pair_score *d_scores;
  c2g_query_result *d_results;
  void *d_fin_head, *d_fin_cand;  // per query / per candidate pose state between the finish kernels (query.cu)
  cudaStream_t qstream[C2G_QUERY_STREAMS];
  cudaEvent_t ev_qfork, ev_qjoin[C2G_QUERY_STREAMS];
  int *d_survivors, *d_nsurv;  // hint slots that pass the thread-per-hint prefilter, and their count
  uint32_t *d_pair_scratch;    // per (query, pre-selected candidate): ellipse pairs of the GMM-L2 refinement (refine.cu)
  int pair_cap;
  long long n_hint_slots;  // max_batch * n_q_levels * C2G_MAX_PIV * nnk
  // staging RING of mirror patches (records + block descriptors), pinned host / device: a patch takes the next free stretch, so the
  // host never waits for the GPU unless the ring wraps onto a patch that is still in flight (the windowed online loop issues ~50
  // small patches per window while the stream is busy with the next window's ingest)
  void *h_patch, *d_patch;
  size_t patch_cap, patch_off;
  struct {
    size_t beg, end;
    cudaEvent_t ev;            // recorded after the kernels that consume the stretch
    int used;
  } patch_ring[C2G_PATCH_RING];
  int patch_head;              // next ring entry to (re)use
  C2gHostDB *hostdb;       // ContourDB::layer_db_ bookkeeping on the host
  int db_dirty;            // device mirror older than the host state
  int db_not_kd;           // some buckets of the mirror are in tree order (fine for the online loop, slow for big batches)
  // windowed online loop (c2g_online_stage / c2g_online_commit): staged = ingested, keys on their way to the host, not yet in the DB
  struct {
    int first_slot, W;
    cudaEvent_t ev_keys;   // the keys of the window have arrived in h_keys
    float *h_keys;         // pinned, [max_batch][C2G_NLEV][C2G_MAX_PIV][C2G_KEY_DIM]
  } staged[2];
  int n_staged, staged_head;  // FIFO of at most two windows (the next one is ingested while the current one is queried)
  unsigned long long *d_work;  // [C2G_WORK_N] work counters of the query kernels (c2g_work_counters); handed to the kernels only while enabled
  int count_work;
  double online_host_s[4];    // host seconds spent by c2g_online_commit: LayerDB bookkeeping, kNN launches, mirror patches, chain launches
  long long online_runs;      // runs of scans that saw identical trees in the windowed loop so far
  long long online_groups;    // kNN launches of the windowed loop so far (a launch serves a group of runs)
  // optional per-kernel timing of the query path (c2g_query_profile): event k is recorded after kernel k - 1
  int prof_on;
  cudaEvent_t prof_ev[C2G_QPROF_N + 1];
  // developer aid (environment C2G_TRACE=<file>): timed events on both streams of the windowed loop, dumped by c2g_destroy
  int trace_on, trace_n;
  struct {
    const char *what;
    cudaEvent_t ev;
    double host_s;
  } trace[C2G_TRACE_CAP];
};

// records a timed event on `st` (no-op unless tracing)
static inline void c2g_trace_mark(c2g_ctx *ctx, const char *what, cudaStream_t st) {
  if (!ctx->trace_on || ctx->trace_n >= C2G_TRACE_CAP) return;
  auto &t = ctx->trace[ctx->trace_n];
  if (cudaEventCreate(&t.ev) != cudaSuccess) return;
  cudaEventRecord(t.ev, st);
  t.what = what;
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  t.host_s = (double) ts.tv_sec + 1e-9 * (double) ts.tv_nsec;
  ctx->trace_n++;
}
