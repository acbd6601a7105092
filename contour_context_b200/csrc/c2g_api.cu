// c2g_api.cu — context + the extern "C" entry points declared in include/c2g.h.
#include <cstdio>
#include <cstring>
#include <cstddef>
#include <new>
#include <vector>

#include "../../include/c2g.h"
#include "c2g_common.cuh"
#include "c2g_ctx.cuh"
#include "c2g_libm.cuh"
#include "layer_db_host.h"
#include "stdsort.cuh"

// launchers implemented in the kernel translation units
int c2g_launch_bev_scatter(const float *pts_dev, const long long *offsets_dev, int B, const C2gIngestParams &P, const C2gBevOut &out,
                           int full_tile, int xyz, int *work_counter, int num_sms, cudaStream_t stream);
int c2g_launch_bev_fill(const c2g_cellkey *tile1, const float *pts_dev, const long long *offsets_dev, int b, int fpp, const C2gIngestParams &P,
                        float *bev_h, float *bev_rf, float *bev_cf, cudaStream_t stream);
int c2g_launch_contours(const C2gBevOut &bev, int B, const C2gIngestParams &P, const int *int_ids_dev, int first_slot,
                        c2g_view *presort_scratch, c2g_scan_head *heads, c2g_view *views, c2g_ell *ells, unsigned char *k2_scratch,
                        int *work_counter, int num_sms, cudaStream_t stream, long long *dbg);
int c2g_launch_warp_sort_selftest(uint32_t *words_dev, int n, int desc, cudaStream_t stream);  // contours.cu
int c2g_contour_max_ctas(int num_sms);
size_t c2g_contour_scratch_bytes(int num_sms, int n_cells, int n_row);
int c2g_query_alloc(c2g_ctx *ctx);
int c2g_db_sync_mode(c2g_ctx *ctx, int want_kd);  // query.cu
void c2g_query_free(c2g_ctx *ctx);
int c2g_online_commit_impl(c2g_ctx *ctx, int first_slot, int W, const float *keys_host, const double *ts_host, const int *seeds_host,
                           const c2g_score_ensemble *lb, const c2g_score_ensemble *ub, c2g_query_result *results_host);  // query.cu
int c2g_refine_alloc(c2g_ctx *ctx);  // refine.cu
void c2g_refine_free(c2g_ctx *ctx);

namespace {

int make_params(const c2g_cm_config &cfg, C2gIngestParams &P) {
  if (cfg.n_levels != C2G_NLEV) return C2G_ERR_ARG;
  if (cfg.n_row <= 0 || cfg.n_col <= 0 || cfg.n_row * cfg.n_col > C2G_MAX_CELLS) return C2G_ERR_ARG;
  if (cfg.n_row % 2 || cfg.n_col % 2) return C2G_ERR_ARG;  // CHECK in contour_mng.h:479-480
  if (cfg.n_row > 255 || cfg.n_col > 255) return C2G_ERR_ARG;
  if (cfg.n_row * ((cfg.n_col + 31) / 32) > 800) return C2G_ERR_ARG;  // bit-plane capacity of the contour kernel (150 x 5 = 750)
  // a level-(l+1) contour must lie inside a level-l contour (the recursion of makeContourRecursiveHelper descends into the
  // parent's ROI): the thresholds have to increase, as in both shipped configurations
  for (int l = 1; l < C2G_NLEV; ++l)
    if (!(cfg.lv_grads[l] > cfg.lv_grads[l - 1])) return C2G_ERR_ARG;
  if (cfg.piv_firsts < 0 || cfg.piv_firsts > C2G_MAX_PIV) return C2G_ERR_ARG;
  if (cfg.dist_firsts < 0 || cfg.dist_firsts > C2G_MAX_DIST_FIRSTS) return C2G_ERR_ARG;
  if (!(cfg.roi_radius > 0.0f) || cfg.roi_radius > 10.0f) return C2G_ERR_ARG;  // key window list capacity
  P.cfg = cfg;
  // ContourManager ctor (contour_mng.h:483-486) + hashPointToImage's padding (contour_mng.h:450), all in float
  const float x_min = -(float) (cfg.n_row / 2) * cfg.reso_row, x_max = -x_min;
  const float y_min = -(float) (cfg.n_col / 2) * cfg.reso_col, y_max = -y_min;
  const float padding = 1e-2f;
  P.x_min_pad = x_min + padding;
  P.x_max_pad = x_max - padding;
  P.y_min_pad = y_min + padding;
  P.y_max_pad = y_max - padding;
  P.half_row = cfg.n_row / 2;
  P.half_col = cfg.n_col / 2;
  P.half_row_f = (float) P.half_row;
  P.half_col_f = (float) P.half_col;
  P.n_cells = cfg.n_row * cfg.n_col;
  return 0;
}

// Which exp() does this host's libm implement? (x86-64 glibc dispatches to an FMA build on CPUs with FMA; the two differ in
// ~0.07 % of results.)  The device then runs the same variant, so GPU keys match what the reference computes on THIS host.
const uint64_t kExpTabHost[256] = C2G_EXP_TAB_INIT;
int probe_exp_mode() {
  int bad1 = 0, bad2 = 0;
  uint64_t st = 0x9E3779B97F4A7C15ull;
  for (int i = 0; i < 400000; ++i) {
    st ^= st << 13;
    st ^= st >> 7;
    st ^= st << 17;
    const double u = (double) (st >> 11) * (1.0 / 9007199254740992.0);
    double x;
    if (i & 1) {
      const float t = (float) (u * 12.0);  // gaussPDF arguments: -0.5 * t * t with t a float difference
      x = (-0.5 * (double) t) * (double) t;
    } else
      x = -70.0 * u;
    const double ref = exp(x);
    if (c2g_d2u(c2g_exp_glibc<false>(x, kExpTabHost)) != c2g_d2u(ref)) ++bad1;
    if (c2g_d2u(c2g_exp_glibc<true>(x, kExpTabHost)) != c2g_d2u(ref)) ++bad2;
  }
  if (bad2 == 0) return 2;
  if (bad1 == 0) return 1;
  return 0;
}

// the scatter -> contour hand-off buffers of scans b0.. of the current batch
C2gBevOut bev_out(const c2g_ctx *ctx, int b0) {
  const size_t nwords = (size_t) ctx->P.cfg.n_row * ((ctx->P.cfg.n_col + 31) / 32);
  C2gBevOut o;
  o.planes = ctx->d_planes + (size_t) b0 * C2G_NLEV * nwords;
  o.fg = ctx->d_fg + (size_t) b0 * ctx->P.n_cells;
  o.hdr = ctx->d_hdr + b0;
  o.tiles = nullptr;
  return o;
}

int stage_inputs(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device, int fpp, const float **pts_dev) {
  if (!ctx || !pts || !offsets_host || B <= 0 || B > ctx->max_batch) return C2G_ERR_ARG;
  const long long total = offsets_host[B] - offsets_host[0];
  if (total < 0) return C2G_ERR_ARG;
  C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_offsets, offsets_host, sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, ctx->stream));
  ctx->last_offsets = ctx->d_offsets;
  if (pts_on_device) {
    if (((uintptr_t) pts) & (fpp == 4 ? 15 : 3)) return C2G_ERR_ARG;
    *pts_dev = pts;
  } else {
    if (offsets_host[B] > ctx->max_points) return C2G_ERR_CAPACITY;
    // staging buffer 0 may still be the target of an earlier pipelined ingest's copy stream or be read by its kernels
    C2G_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ctx->ev_stage_free[0], 0));
    C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_pts_stage2[0], pts, sizeof(float) * fpp * (size_t) offsets_host[B], cudaMemcpyHostToDevice, ctx->stream));
    *pts_dev = ctx->d_pts_stage2[0];
  }
  return 0;
}

}  // namespace

extern "C" {

int c2g_abi_version(void) { return 1; }

int c2g_sizeof(int which) {
  switch (which) {
    case 0: return (int) sizeof(c2g_scan_head);
    case 1: return (int) sizeof(c2g_view);
    case 2: return (int) sizeof(c2g_bci);
    case 3: return (int) sizeof(c2g_hint);
    case 4: return (int) sizeof(c2g_pair_score);
    case 5: return (int) sizeof(c2g_query_result);
    case 6: return (int) sizeof(c2g_cm_config);
    case 7: return (int) sizeof(c2g_db_config);
    default: return -1;
  }
}

int c2g_create(const c2g_cm_config *cm_cfg, const c2g_db_config *db_cfg, int device, int scan_capacity, int max_batch,
               long long max_points, c2g_ctx **out) {
  if (!cm_cfg || !db_cfg || !out || scan_capacity <= 0 || max_batch <= 0 || max_points <= 0) return C2G_ERR_ARG;
  if (db_cfg->n_q_levels <= 0 || db_cfg->n_q_levels > C2G_NUM_Q_LEVELS_MAX || db_cfg->nnk <= 0 || db_cfg->nnk > 64) return C2G_ERR_ARG;
  for (int i = 0; i < db_cfg->n_q_levels; ++i)
    if (db_cfg->q_levels[i] < 1 || db_cfg->q_levels[i] > 4) return C2G_ERR_ARG;  // pair bitmap covers levels 1..4
  c2g_ctx *ctx = new (std::nothrow) c2g_ctx();
  if (!ctx) return C2G_ERR_ARG;
  memset(ctx, 0, sizeof(*ctx));
  int rc = make_params(*cm_cfg, ctx->P);
  if (rc) {
    delete ctx;
    return rc;
  }
  ctx->P.exp_mode = probe_exp_mode();
  ctx->trace_on = getenv("C2G_TRACE") != nullptr;
  ctx->db = *db_cfg;
  ctx->device = device;
  ctx->scan_cap = scan_capacity;
  ctx->max_batch = max_batch;
  ctx->max_points = max_points;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e == cudaSuccess && (device < 0 || device >= n_dev)) e = cudaErrorInvalidDevice;
  if (e != cudaSuccess) {
    delete ctx;
    return -(int) e;
  }
  C2gDeviceGuard guard(device);  // the caller's current device is restored on every exit path
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) {
    delete ctx;
    return -(int) e;
  }
  ctx->num_sms = prop.multiProcessorCount;
  const size_t ncell = ctx->P.n_cells;
#define ALLOC(ptr, bytes)                                        \
  do {                                                           \
    e = cudaMalloc((void **) &(ptr), (bytes));                   \
    if (e != cudaSuccess) {                                      \
      c2g_destroy(ctx);                                          \
      return -(int) e;                                           \
    }                                                            \
  } while (0)
  e = cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete ctx;
    return -(int) e;
  }
  ctx->stream = ctx->own_stream;
  ALLOC(ctx->d_pts_stage2[0], sizeof(float) * 4 * (size_t) max_points);
  ALLOC(ctx->d_pts_stage2[1], sizeof(float) * 4 * (size_t) max_points);
  e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    c2g_destroy(ctx);
    return -(int) e;
  }
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_stage_free[i], cudaEventDisableTiming);
  for (int i = 0; i < C2G_MAX_CHUNK_EVENTS && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&ctx->ev_chunk[i], cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_patch_up, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    c2g_destroy(ctx);
    return -(int) e;
  }
  ALLOC(ctx->d_offsets, sizeof(long long) * (max_batch + 1));
  ALLOC(ctx->d_int_ids, sizeof(int) * max_batch);
  for (int i = 0; i < 2; ++i) {
    ALLOC(ctx->d_offsets2[i], sizeof(long long) * (max_batch + 1));
    ALLOC(ctx->d_int_ids2[i], sizeof(int) * max_batch);
  }
  {
    const size_t nwords = (size_t) ctx->P.cfg.n_row * ((ctx->P.cfg.n_col + 31) / 32);
    ALLOC(ctx->d_planes, sizeof(uint32_t) * C2G_NLEV * nwords * max_batch);
    ALLOC(ctx->d_fg, sizeof(float4) * ncell * max_batch);
    ALLOC(ctx->d_hdr, sizeof(int2) * max_batch);
    ALLOC(ctx->d_work_counter_k1, sizeof(int) * (4 + (size_t) max_batch));
    ALLOC(ctx->d_tile1, sizeof(c2g_cellkey) * ncell);
    ALLOC(ctx->d_planes1, sizeof(uint32_t) * C2G_NLEV * nwords);
    ALLOC(ctx->d_fg1, sizeof(float4) * ncell);
    ALLOC(ctx->d_hdr1, sizeof(int2));
  }
  ALLOC(ctx->d_bev_h, sizeof(float) * ncell);
  ALLOC(ctx->d_bev_rf, sizeof(float) * ncell);
  ALLOC(ctx->d_bev_cf, sizeof(float) * ncell);
  ALLOC(ctx->d_presort, sizeof(c2g_view) * C2G_VIEW_CAP * (size_t) c2g_contour_max_ctas(ctx->num_sms));
  ALLOC(ctx->d_heads, sizeof(c2g_scan_head) * (size_t) scan_capacity);
  ALLOC(ctx->d_views, sizeof(c2g_view) * C2G_VIEW_CAP * (size_t) scan_capacity);
  ALLOC(ctx->d_ells, sizeof(c2g_ell) * C2G_VIEW_CAP * (size_t) scan_capacity);
  ALLOC(ctx->d_dbg, sizeof(long long) * 64);
  ALLOC(ctx->d_work_counter, sizeof(int));
  {
    const size_t k2b = c2g_contour_scratch_bytes(ctx->num_sms, ctx->P.n_cells, ctx->P.cfg.n_row);
    ALLOC(ctx->d_k2_scratch, k2b);
    e = cudaMemset(ctx->d_k2_scratch, 0, k2b);  // the arena locks start free
    if (e != cudaSuccess) {
      c2g_destroy(ctx);
      return -(int) e;
    }
  }
#undef ALLOC
  rc = c2g_query_alloc(ctx);
  if (!rc) rc = c2g_refine_alloc(ctx);
  if (rc) {
    c2g_destroy(ctx);
    return rc;
  }
  for (int i = 0; i < 2; ++i) {
    e = cudaHostAlloc((void **) &ctx->staged[i].h_keys, sizeof(float) * C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM * (size_t) max_batch, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->staged[i].ev_keys, cudaEventDisableTiming);
    if (e != cudaSuccess) {
      c2g_destroy(ctx);
      return -(int) e;
    }
  }
  ctx->hostdb = new (std::nothrow) C2gHostDB();
  if (!ctx->hostdb) {
    c2g_destroy(ctx);
    return C2G_ERR_CAPACITY;
  }
  c2g_hostdb_init(*ctx->hostdb, db_cfg->n_q_levels, db_cfg->max_elapse, db_cfg->min_elapse);
  ctx->db_dirty = 0;
  ctx->db_not_kd = 0;
  *out = ctx;
  return 0;
}

int c2g_destroy(c2g_ctx *ctx) {
  if (!ctx) return 0;
  C2gDeviceGuard guard(ctx->device);
  cudaDeviceSynchronize();
  if (ctx->trace_on && ctx->trace_n > 0) {  // developer aid: device and host time of every mark, ms since the first one
    if (FILE *f = fopen(getenv("C2G_TRACE"), "a")) {
      fprintf(f, "# context on device %d: %d marks (what, device ms, host ms)\n", ctx->device, ctx->trace_n);
      for (int i = 0; i < ctx->trace_n; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->trace[0].ev, ctx->trace[i].ev);
        fprintf(f, "%s %.3f %.3f\n", ctx->trace[i].what, ms, (ctx->trace[i].host_s - ctx->trace[0].host_s) * 1e3);
      }
      fclose(f);
    }
    for (int i = 0; i < ctx->trace_n; ++i) cudaEventDestroy(ctx->trace[i].ev);
  }
  c2g_query_free(ctx);
  c2g_refine_free(ctx);
  delete ctx->hostdb;
  for (int i = 0; i < 2; ++i) {
    if (ctx->staged[i].h_keys) cudaFreeHost(ctx->staged[i].h_keys);
    if (ctx->staged[i].ev_keys) cudaEventDestroy(ctx->staged[i].ev_keys);
  }
  cudaFree(ctx->d_pts_stage2[0]);
  cudaFree(ctx->d_pts_stage2[1]);
  for (int i = 0; i < 2; ++i)
    if (ctx->ev_stage_free[i]) cudaEventDestroy(ctx->ev_stage_free[i]);
  for (int i = 0; i < C2G_MAX_CHUNK_EVENTS; ++i)
    if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  cudaFree(ctx->d_offsets);
  cudaFree(ctx->d_int_ids);
  for (int i = 0; i < 2; ++i) {
    cudaFree(ctx->d_offsets2[i]);
    cudaFree(ctx->d_int_ids2[i]);
  }
  if (ctx->ev_patch_up) cudaEventDestroy(ctx->ev_patch_up);
  cudaFree(ctx->d_planes);
  cudaFree(ctx->d_fg);
  cudaFree(ctx->d_hdr);
  cudaFree(ctx->d_work_counter_k1);
  cudaFree(ctx->d_tile1);
  cudaFree(ctx->d_planes1);
  cudaFree(ctx->d_fg1);
  cudaFree(ctx->d_hdr1);
  cudaFree(ctx->d_bev_h);
  cudaFree(ctx->d_bev_rf);
  cudaFree(ctx->d_bev_cf);
  cudaFree(ctx->d_presort);
  cudaFree(ctx->d_heads);
  cudaFree(ctx->d_views);
  cudaFree(ctx->d_ells);
  cudaFree(ctx->d_dbg);
  cudaFree(ctx->d_work_counter);
  cudaFree(ctx->d_k2_scratch);
  if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return 0;
}

int c2g_host_alloc(void **out, size_t bytes) {
  if (!out || bytes == 0) return C2G_ERR_ARG;
  C2G_CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 0;
}

int c2g_host_free(void *p) {
  if (p) C2G_CUDA_TRY(cudaFreeHost(p));
  return 0;
}

int c2g_set_stream(c2g_ctx *ctx, void *cuda_stream) {
  if (!ctx) return C2G_ERR_ARG;
  ctx->stream = cuda_stream ? (cudaStream_t) cuda_stream : ctx->own_stream;
  return 0;
}

int c2g_sync(c2g_ctx *ctx) {
  if (!ctx) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

static int bev_only_impl(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device, int fpp) {
  if (!ctx) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  const float *pts_dev = nullptr;
  int rc = stage_inputs(ctx, pts, offsets_host, B, pts_on_device, fpp, &pts_dev);
  if (rc) return rc;
  rc = c2g_launch_bev_scatter(pts_dev, ctx->d_offsets, B, ctx->P, bev_out(ctx, 0), 0, fpp == 3, ctx->d_work_counter_k1, ctx->num_sms, ctx->stream);
  if (rc) return rc;
  // the staging buffer is free again once the kernel that reads it has run (a later c2g_get_bev of this batch is ordered
  // behind it on the same stream; a later pipelined ingest waits for this event before its copy stream overwrites the buffer)
  if (!pts_on_device) C2G_CUDA_TRY(cudaEventRecord(ctx->ev_stage_free[0], ctx->stream));
  ctx->launches += 1;
  ctx->last_B = B;
  ctx->last_pts = pts_dev;
  ctx->last_fpp = fpp;
  return 0;
}

int c2g_ingest_bev_only(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device) {
  return bev_only_impl(ctx, pts, offsets_host, B, pts_on_device, 4);
}

// Host inputs: the batch is cut into chunks of one wave (num_sms scans); chunk k+1 crosses PCIe on the copy stream while the
// kernels of chunk k run, and the two staging buffers alternate between calls so that the copy of the NEXT call starts while
// this call's query kernels are still running.  Device inputs: two launches for the whole batch.
static int ingest_host_pipelined(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int first_slot, const int *int_ids_host, int fpp) {
  if (offsets_host[B] - offsets_host[0] > ctx->max_points) return C2G_ERR_CAPACITY;
  const int cur = ctx->stage_sel;
  ctx->stage_sel ^= 1;
  float *stage = ctx->d_pts_stage2[cur];
  // offsets relative to the staging buffer
  std::vector<long long> rel((size_t) B + 1);
  for (int i = 0; i <= B; ++i) rel[i] = offsets_host[i] - offsets_host[0];
  // the copy stream may overwrite this staging buffer (points, offsets, ids) only after the kernels that last read it have
  // finished.  Everything host->device goes through the copy stream: a small copy issued on the kernel stream would sit in
  // the copy queue behind whatever that stream still has to run, in front of the next batch's points.
  C2G_CUDA_TRY(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_stage_free[cur], 0));
  long long *d_off = ctx->d_offsets2[cur];
  C2G_CUDA_TRY(cudaMemcpyAsync(d_off, rel.data(), sizeof(long long) * (B + 1), cudaMemcpyHostToDevice, ctx->copy_stream));
  const int *ids_dev = nullptr;
  if (int_ids_host) {
    C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_int_ids2[cur], int_ids_host, sizeof(int) * B, cudaMemcpyHostToDevice, ctx->copy_stream));
    ids_dev = ctx->d_int_ids2[cur];
  }
  ctx->last_offsets = d_off;
  const int CH = ctx->num_sms;
  int k = 0;
  for (int b0 = 0; b0 < B; b0 += CH, ++k) {
    const int n = (B - b0 < CH) ? (B - b0) : CH;
    const size_t p0 = (size_t) rel[b0], p1 = (size_t) rel[b0 + n];
    c2g_trace_mark(ctx, "h2d_begin", ctx->copy_stream);
    C2G_CUDA_TRY(cudaMemcpyAsync(stage + fpp * p0, pts + fpp * ((size_t) offsets_host[0] + p0), sizeof(float) * fpp * (p1 - p0), cudaMemcpyHostToDevice, ctx->copy_stream));
    c2g_trace_mark(ctx, "h2d_end", ctx->copy_stream);
    cudaEvent_t ev = ctx->ev_chunk[k % C2G_MAX_CHUNK_EVENTS];
    C2G_CUDA_TRY(cudaEventRecord(ev, ctx->copy_stream));
    C2G_CUDA_TRY(cudaStreamWaitEvent(ctx->stream, ev, 0));
    const C2gBevOut bo = bev_out(ctx, b0);
    c2g_trace_mark(ctx, "ingest_begin", ctx->stream);
    int rc = c2g_launch_bev_scatter(stage, d_off + b0, n, ctx->P, bo, 0, fpp == 3, ctx->d_work_counter_k1, ctx->num_sms, ctx->stream);
    if (rc) return rc;
    rc = c2g_launch_contours(bo, n, ctx->P, ids_dev ? ids_dev + b0 : nullptr, first_slot + b0, ctx->d_presort, ctx->d_heads, ctx->d_views, ctx->d_ells,
                             ctx->d_k2_scratch, ctx->d_work_counter, ctx->num_sms, ctx->stream, ctx->d_dbg);
    if (rc) return rc;
    ctx->launches += 2;
    c2g_trace_mark(ctx, "ingest_end", ctx->stream);
  }
  C2G_CUDA_TRY(cudaEventRecord(ctx->ev_stage_free[cur], ctx->stream));
  ctx->last_B = B;
  ctx->last_pts = stage;
  ctx->last_fpp = fpp;
  return 0;
}

static int ingest_impl(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device, int first_slot,
                       const int *int_ids_host, int fpp) {
  if (!ctx || !pts || !offsets_host || B <= 0 || B > ctx->max_batch || first_slot < 0 || first_slot + B > ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  if (!pts_on_device) return ingest_host_pipelined(ctx, pts, offsets_host, B, first_slot, int_ids_host, fpp);
  const int *ids_dev = nullptr;
  if (int_ids_host) {
    C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_int_ids, int_ids_host, sizeof(int) * B, cudaMemcpyHostToDevice, ctx->stream));
    ids_dev = ctx->d_int_ids;
  }
  int rc = bev_only_impl(ctx, pts, offsets_host, B, pts_on_device, fpp);
  if (rc) return rc;
  rc = c2g_launch_contours(bev_out(ctx, 0), B, ctx->P, ids_dev, first_slot, ctx->d_presort, ctx->d_heads, ctx->d_views, ctx->d_ells, ctx->d_k2_scratch,
                           ctx->d_work_counter, ctx->num_sms, ctx->stream, ctx->d_dbg);
  if (rc) return rc;
  ctx->launches += 1;
  return 0;
}

int c2g_ingest(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int B, int pts_on_device, int first_slot,
               const int *int_ids_host) {
  return ingest_impl(ctx, pts, offsets_host, B, pts_on_device, first_slot, int_ids_host, 4);
}

int c2g_ingest_xyz(c2g_ctx *ctx, const float *xyz, const long long *offsets_host, int B, int pts_on_device, int first_slot,
                   const int *int_ids_host) {
  return ingest_impl(ctx, xyz, offsets_host, B, pts_on_device, first_slot, int_ids_host, 3);
}

int c2g_get_heads(c2g_ctx *ctx, int first_slot, int n, c2g_scan_head *out_host) {
  if (!ctx || !out_host || first_slot < 0 || n < 0 || first_slot + n > ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaMemcpyAsync(out_host, ctx->d_heads + first_slot, sizeof(c2g_scan_head) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_get_views(c2g_ctx *ctx, int slot, c2g_view *out_host) {
  if (!ctx || !out_host || slot < 0 || slot >= ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaMemcpyAsync(out_host, ctx->d_views + (size_t) slot * C2G_VIEW_CAP, sizeof(c2g_view) * C2G_VIEW_CAP, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// The ingest kernels keep only what contours, moments and keys read (bit-planes + the foreground cells).  The dense image
// (getBevImage, bev_pixfs_) and the raw 64-bit cells are produced on demand: the full-tile variant of the scatter kernel runs
// again on the one requested scan of the last batch (its points are still resident: staging buffer or caller's buffer).
static int rescatter_full(c2g_ctx *ctx, int batch_index) {
  if (!ctx->last_pts) return C2G_ERR_STATE;
  C2gBevOut o;
  o.planes = ctx->d_planes1;
  o.fg = ctx->d_fg1;
  o.hdr = ctx->d_hdr1;
  o.tiles = ctx->d_tile1;
  int rc = c2g_launch_bev_scatter(ctx->last_pts, ctx->last_offsets + batch_index, 1, ctx->P, o, 1, ctx->last_fpp == 3, ctx->d_work_counter_k1, ctx->num_sms,
                                  ctx->stream);
  if (rc) return rc;
  ctx->launches += 1;
  return 0;
}

int c2g_get_bev(c2g_ctx *ctx, int batch_index, float *bev, float *row_f, float *col_f) {
  if (!ctx || batch_index < 0 || batch_index >= ctx->last_B) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  const size_t n = ctx->P.n_cells;
  int rc = rescatter_full(ctx, batch_index);
  if (rc) return rc;
  rc = c2g_launch_bev_fill(ctx->d_tile1, ctx->last_pts, ctx->last_offsets, batch_index, ctx->last_fpp, ctx->P, ctx->d_bev_h, ctx->d_bev_rf, ctx->d_bev_cf,
                           ctx->stream);
  if (rc) return rc;
  ctx->launches += 1;
  if (bev) C2G_CUDA_TRY(cudaMemcpyAsync(bev, ctx->d_bev_h, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (row_f) C2G_CUDA_TRY(cudaMemcpyAsync(row_f, ctx->d_bev_rf, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  if (col_f) C2G_CUDA_TRY(cudaMemcpyAsync(col_f, ctx->d_bev_cf, sizeof(float) * n, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_get_tiles(c2g_ctx *ctx, int batch_index, unsigned long long *out_host) {
  if (!ctx || !out_host || batch_index < 0 || batch_index >= ctx->last_B) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  int rc = rescatter_full(ctx, batch_index);
  if (rc) return rc;
  C2G_CUDA_TRY(cudaMemcpyAsync(out_host, ctx->d_tile1, sizeof(c2g_cellkey) * (size_t) ctx->P.n_cells, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_get_bev_compact(c2g_ctx *ctx, int batch_index, int full_tile_variant, unsigned int *planes_host, float *fg_host, int *hdr_host) {
  if (!ctx || batch_index < 0 || batch_index >= ctx->last_B) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  const size_t nwords = (size_t) ctx->P.cfg.n_row * ((ctx->P.cfg.n_col + 31) / 32), ncell = ctx->P.n_cells;
  C2gBevOut o = bev_out(ctx, batch_index);
  if (full_tile_variant) {
    int rc = rescatter_full(ctx, batch_index);
    if (rc) return rc;
    o.planes = ctx->d_planes1;
    o.fg = ctx->d_fg1;
    o.hdr = ctx->d_hdr1;
  }
  int hdr[2] = {0, 0};
  C2G_CUDA_TRY(cudaMemcpyAsync(hdr, o.hdr, sizeof(int2), cudaMemcpyDeviceToHost, ctx->stream));
  if (planes_host) C2G_CUDA_TRY(cudaMemcpyAsync(planes_host, o.planes, sizeof(uint32_t) * C2G_NLEV * nwords, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (hdr[1] < 0 || (size_t) hdr[1] > ncell) return C2G_ERR_STATE;
  if (fg_host && hdr[1] > 0) {
    C2G_CUDA_TRY(cudaMemcpyAsync(fg_host, o.fg, sizeof(float4) * (size_t) hdr[1], cudaMemcpyDeviceToHost, ctx->stream));
    C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  }
  if (hdr_host) {
    hdr_host[0] = hdr[0];
    hdr_host[1] = hdr[1];
  }
  return 0;
}

int c2g_copy_slots(c2g_ctx *ctx, int src_first, int dst_first, int n) {
  if (!ctx || n < 0 || src_first < 0 || dst_first < 0 || src_first + n > ctx->scan_cap || dst_first + n > ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  if (n == 0 || src_first == dst_first) return 0;
  C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_heads + dst_first, ctx->d_heads + src_first, sizeof(c2g_scan_head) * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
  C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_views + (size_t) dst_first * C2G_VIEW_CAP, ctx->d_views + (size_t) src_first * C2G_VIEW_CAP,
                               sizeof(c2g_view) * C2G_VIEW_CAP * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
  C2G_CUDA_TRY(cudaMemcpyAsync(ctx->d_ells + (size_t) dst_first * C2G_VIEW_CAP, ctx->d_ells + (size_t) src_first * C2G_VIEW_CAP,
                               sizeof(c2g_ell) * C2G_VIEW_CAP * (size_t) n, cudaMemcpyDeviceToDevice, ctx->stream));
  return 0;
}

int c2g_db_add_scans(c2g_ctx *ctx, int first_slot, int n, const double *ts_host) {
  if (!ctx || !ts_host || n <= 0 || first_slot < 0 || first_slot + n > ctx->scan_cap) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2gHostDB &db = *ctx->hostdb;
  if (first_slot != db.n_scans || ctx->n_staged) return C2G_ERR_STATE;  // gidx == all_bevs_.size() at the time of addScan
  std::vector<float> keys((size_t) n * C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM);
  const size_t kbytes = sizeof(float) * C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM;
  C2G_CUDA_TRY(cudaMemcpy2DAsync(keys.data(), kbytes, (const char *) (ctx->d_heads + first_slot) + offsetof(c2g_scan_head, keys),
                                 sizeof(c2g_scan_head), kbytes, (size_t) n, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) {
    const float *sk = keys.data() + (size_t) i * C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM;
    for (int ll = 0; ll < ctx->db.n_q_levels; ++ll) {
      const int lev = ctx->db.q_levels[ll];
      for (int seq = 0; seq < ctx->P.cfg.piv_firsts; ++seq)
        c2g_hostdb_push(db, ll, sk + ((size_t) lev * C2G_MAX_PIV + seq) * C2G_KEY_DIM, ts_host[i], first_slot + i, seq);
    }
    db.n_scans++;
  }
  ctx->db_dirty = 1;
  return 0;
}

static int online_stage_impl(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host, int fpp) {
  if (!ctx || W <= 0 || W > ctx->max_batch) return C2G_ERR_ARG;
  if (ctx->n_staged >= 2) return C2G_ERR_STATE;
  int first_slot = ctx->hostdb->n_scans;
  for (int k = 0; k < ctx->n_staged; ++k) first_slot += ctx->staged[(ctx->staged_head + k) & 1].W;
  if (first_slot + W > ctx->scan_cap) return C2G_ERR_CAPACITY;
  int rc = ingest_impl(ctx, pts, offsets_host, W, pts_on_device, first_slot, int_ids_host, fpp);
  if (rc) return rc;
  C2gDeviceGuard guard(ctx->device);
  auto &st = ctx->staged[(ctx->staged_head + ctx->n_staged) & 1];
  st.first_slot = first_slot;
  st.W = W;
  const size_t kbytes = sizeof(float) * C2G_NLEV * C2G_MAX_PIV * C2G_KEY_DIM;
  C2G_CUDA_TRY(cudaMemcpy2DAsync(st.h_keys, kbytes, (const char *) (ctx->d_heads + first_slot) + offsetof(c2g_scan_head, keys), sizeof(c2g_scan_head), kbytes,
                                 (size_t) W, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaEventRecord(st.ev_keys, ctx->stream));
  c2g_trace_mark(ctx, "keys_d2h_end", ctx->stream);
  ctx->n_staged++;
  return 0;
}

int c2g_online_stage(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host) {
  return online_stage_impl(ctx, pts, offsets_host, W, pts_on_device, int_ids_host, 4);
}

int c2g_online_stage_xyz(c2g_ctx *ctx, const float *xyz, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host) {
  return online_stage_impl(ctx, xyz, offsets_host, W, pts_on_device, int_ids_host, 3);
}

int c2g_online_commit(c2g_ctx *ctx, const double *ts_host, const int *seeds_host, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
                      c2g_query_result *results_host) {
  if (!ctx || !ts_host || !seeds_host || !lb || !ub) return C2G_ERR_ARG;
  if (ctx->n_staged <= 0) return C2G_ERR_STATE;
  C2gDeviceGuard guard(ctx->device);
  auto &st = ctx->staged[ctx->staged_head];
  c2g_trace_mark(ctx, "host:commit_enter", ctx->copy_stream);
  C2G_CUDA_TRY(cudaEventSynchronize(st.ev_keys));  // the only wait of the window: its 1.4 KB of keys per scan
  c2g_trace_mark(ctx, "host:keys_arrived", ctx->copy_stream);
  int rc = c2g_online_commit_impl(ctx, st.first_slot, st.W, st.h_keys, ts_host, seeds_host, lb, ub, results_host);
  if (rc) return rc;
  ctx->staged_head ^= 1;
  ctx->n_staged--;
  return 0;
}

int c2g_online_window(c2g_ctx *ctx, const float *pts, const long long *offsets_host, int W, int pts_on_device, const int *int_ids_host,
                      const double *ts_host, const int *seeds_host, const c2g_score_ensemble *lb, const c2g_score_ensemble *ub,
                      c2g_query_result *results_host) {
  if (ctx && ctx->n_staged != 0) return C2G_ERR_STATE;
  int rc = c2g_online_stage(ctx, pts, offsets_host, W, pts_on_device, int_ids_host);
  if (rc) return rc;
  rc = c2g_online_commit(ctx, ts_host, seeds_host, lb, ub, results_host);
  if (rc) return rc;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

long long c2g_online_runs(c2g_ctx *ctx) { return ctx ? ctx->online_runs : 0; }
long long c2g_online_groups(c2g_ctx *ctx) { return ctx ? ctx->online_groups : 0; }

int c2g_online_host_seconds(c2g_ctx *ctx, double *out4) {
  if (!ctx || !out4) return C2G_ERR_ARG;
  for (int k = 0; k < 4; ++k) out4[k] = ctx->online_host_s[k];
  return 0;
}

int c2g_work_counters(c2g_ctx *ctx, int enable, unsigned long long *out_host) {
  if (!ctx) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  if (out_host) C2G_CUDA_TRY(cudaMemcpy(out_host, ctx->d_work, sizeof(unsigned long long) * C2G_WORK_N, cudaMemcpyDeviceToHost));
  C2G_CUDA_TRY(cudaMemset(ctx->d_work, 0, sizeof(unsigned long long) * C2G_WORK_N));
  ctx->count_work = enable ? 1 : 0;
  return 0;
}

int c2g_db_push_and_balance(c2g_ctx *ctx, int seed, double ts) {
  if (!ctx) return C2G_ERR_ARG;
  c2g_hostdb_push_and_balance(*ctx->hostdb, seed, ts);
  ctx->db_dirty = 1;
  return 0;
}

int c2g_db_size(c2g_ctx *ctx) { return ctx ? ctx->hostdb->n_scans : C2G_ERR_ARG; }

int c2g_db_sync(c2g_ctx *ctx) {
  if (!ctx) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  if (!ctx->db_dirty && !ctx->db_not_kd) return 0;
  int rc = c2g_db_sync_mode(ctx, 1);  // query.cu: patches the mirror, kd-blocks every bucket
  if (rc) return rc;
  ctx->db_dirty = 0;
  ctx->db_not_kd = 0;
  return 0;
}

int c2g_db_layer_state(c2g_ctx *ctx, int ll, float *bucket_ranges, int *tree_sizes, int *buffer_sizes) {
  if (!ctx || ll < 0 || ll >= ctx->db.n_q_levels) return C2G_ERR_ARG;
  const C2gLayerHost &L = ctx->hostdb->layers[ll];
  for (int i = 0; i <= C2G_NUM_BUCKETS; ++i) bucket_ranges[i] = L.ranges[i];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
    tree_sizes[i] = (int) L.buckets[i].tree.size();
    buffer_sizes[i] = (int) L.buckets[i].buffer.size();
  }
  return 0;
}

int c2g_db_bucket_tree(c2g_ctx *ctx, int ll, int bucket, float *keys, int *gidx, int *seq) {
  if (!ctx || ll < 0 || ll >= ctx->db.n_q_levels || bucket < 0 || bucket >= C2G_NUM_BUCKETS) return C2G_ERR_ARG;
  const std::vector<C2gKeyRec> &t = ctx->hostdb->layers[ll].buckets[bucket].tree;
  for (size_t i = 0; i < t.size(); ++i) {
    for (int d = 0; d < C2G_KEY_DIM; ++d) keys[i * C2G_KEY_DIM + d] = t[i].k[d];
    gidx[i] = t[i].gidx;
    seq[i] = t[i].seq;
  }
  return 0;
}

/* stand-alone host LayerDB handles (no CUDA context needed): CPU tests of the rebalancing logic */
void *c2g_hostdb_create(int n_layers, double max_elapse, double min_elapse) {
  if (n_layers <= 0 || n_layers > C2G_NUM_Q_LEVELS_MAX) return nullptr;
  C2gHostDB *db = new (std::nothrow) C2gHostDB();
  if (db) c2g_hostdb_init(*db, n_layers, max_elapse, min_elapse);
  return db;
}
void c2g_hostdb_free(void *h) { delete (C2gHostDB *) h; }
int c2g_hostdb_push_key(void *h, int ll, const float *key, double ts, int gidx, int seq) {
  if (!h || !key) return C2G_ERR_ARG;
  c2g_hostdb_push(*(C2gHostDB *) h, ll, key, ts, gidx, seq);
  return 0;
}
int c2g_hostdb_balance(void *h, int seed, double ts) {
  if (!h) return C2G_ERR_ARG;
  c2g_hostdb_push_and_balance(*(C2gHostDB *) h, seed, ts);
  return 0;
}
int c2g_hostdb_state(void *h, int ll, float *bucket_ranges, int *tree_sizes, int *buffer_sizes) {
  if (!h) return C2G_ERR_ARG;
  const C2gLayerHost &L = ((C2gHostDB *) h)->layers[ll];
  for (int i = 0; i <= C2G_NUM_BUCKETS; ++i) bucket_ranges[i] = L.ranges[i];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) {
    tree_sizes[i] = (int) L.buckets[i].tree.size();
    buffer_sizes[i] = (int) L.buckets[i].buffer.size();
  }
  return 0;
}
int c2g_hostdb_indexed(void *h, int ll, int *indexed) {
  if (!h || !indexed) return C2G_ERR_ARG;
  const C2gLayerHost &L = ((C2gHostDB *) h)->layers[ll];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) indexed[i] = (int) L.buckets[i].indexed;
  return 0;
}
int c2g_hostdb_tree(void *h, int ll, int bucket, float *keys, int *gidx, int *seq) {
  if (!h) return C2G_ERR_ARG;
  const std::vector<C2gKeyRec> &t = ((C2gHostDB *) h)->layers[ll].buckets[bucket].tree;
  for (size_t i = 0; i < t.size(); ++i) {
    for (int d = 0; d < C2G_KEY_DIM; ++d) keys[i * C2G_KEY_DIM + d] = t[i].k[d];
    gidx[i] = t[i].gidx;
    seq[i] = t[i].seq;
  }
  return 0;
}

int c2g_hostdb_versions(void *h, int ll, unsigned int *restructured) {
  if (!h || !restructured) return C2G_ERR_ARG;
  const C2gLayerHost &L = ((C2gHostDB *) h)->layers[ll];
  for (int i = 0; i < C2G_NUM_BUCKETS; ++i) restructured[i] = L.buckets[i].restructured;
  return 0;
}

int c2g_exp_mode(c2g_ctx *ctx) { return ctx ? ctx->P.exp_mode : C2G_ERR_ARG; }

/* host execution of the libm restatements in csrc/c2g_libm.cuh (tests): kind 0 exp (glibc, no FMA), 1 exp (glibc, FMA),
 * 2 atan2f (in = y,x pairs), 3 acosf, 4 atanf, 5 = probe_exp_mode() (returns the mode, ignores the buffers) */
int c2g_selftest_libm(int kind, int n, const void *in, void *out) {
  if (kind == 5) return probe_exp_mode();
  if (!in || !out || n < 0) return C2G_ERR_ARG;
  for (int i = 0; i < n; ++i) {
    switch (kind) {
      case 0: ((double *) out)[i] = c2g_exp_glibc<false>(((const double *) in)[i], kExpTabHost); break;
      case 1: ((double *) out)[i] = c2g_exp_glibc<true>(((const double *) in)[i], kExpTabHost); break;
      case 2: ((float *) out)[i] = c2g_atan2f(((const float *) in)[2 * i], ((const float *) in)[2 * i + 1]); break;
      case 3: ((float *) out)[i] = c2g_acosf(((const float *) in)[i]); break;
      case 4: ((float *) out)[i] = c2g_atanf(((const float *) in)[i]); break;
      default: return C2G_ERR_ARG;
    }
  }
  return 0;
}

long long c2g_launch_count(c2g_ctx *ctx) { return ctx ? ctx->launches : 0; }

int c2g_query_profile(c2g_ctx *ctx, int enable, float *ms_out) {
  if (!ctx) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  if (ms_out) {  // durations of the last profiled c2g_query_async: knn, prefilter, score, replay, corr, output, refine, rank
    if (!ctx->prof_on) return C2G_ERR_STATE;
    C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < 8; ++k) C2G_CUDA_TRY(cudaEventElapsedTime(&ms_out[k], ctx->prof_ev[k], ctx->prof_ev[k + 1]));
  }
  if (enable && !ctx->prof_on) {
    for (int k = 0; k <= C2G_QPROF_N; ++k) C2G_CUDA_TRY(cudaEventCreate(&ctx->prof_ev[k]));
    ctx->prof_on = 1;
  } else if (!enable && ctx->prof_on) {
    for (int k = 0; k <= C2G_QPROF_N; ++k) cudaEventDestroy(ctx->prof_ev[k]);
    ctx->prof_on = 0;
  }
  return 0;
}

/* developer aid: per-phase clock64() stamps of CTA 0's first scan in the last contour kernel (64 values) */
int c2g_scatter_deferred(c2g_ctx *ctx, int *n_out) {
  if (!ctx || !n_out) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaMemcpyAsync(n_out, ctx->d_work_counter_k1 + 2, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_debug_clocks(c2g_ctx *ctx, long long *out_host) {
  if (!ctx || !out_host) return C2G_ERR_ARG;
  C2gDeviceGuard guard(ctx->device);
  C2G_CUDA_TRY(cudaMemcpyAsync(out_host, ctx->d_dbg, sizeof(long long) * 64, cudaMemcpyDeviceToHost, ctx->stream));
  C2G_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int c2g_selftest_warpsort(c2g_ctx *ctx, unsigned int *words, int n, int desc) {
  if (!ctx || !words || n < 0 || n > 20000) return C2G_ERR_ARG;
  if (n == 0) return 0;
  C2gDeviceGuard guard(ctx->device);
  uint32_t *d = nullptr;
  C2G_CUDA_TRY(cudaMalloc((void **) &d, sizeof(uint32_t) * (size_t) n));
  cudaError_t e = cudaMemcpyAsync(d, words, sizeof(uint32_t) * (size_t) n, cudaMemcpyHostToDevice, ctx->stream);
  int rc = e == cudaSuccess ? c2g_launch_warp_sort_selftest(d, n, desc, ctx->stream) : -(int) e;
  if (!rc) {
    e = cudaMemcpyAsync(words, d, sizeof(uint32_t) * (size_t) n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) rc = -(int) e;
  }
  cudaFree(d);
  return rc;
}

int c2g_selftest_stdsort(unsigned int *words, int n, int desc) {
  if (!words || n < 0) return C2G_ERR_ARG;
  if (desc)
    c2g_sort::std_sort(words, (long) n, [](unsigned a, unsigned b) { return (a >> 16) > (b >> 16); });
  else
    c2g_sort::std_sort(words, (long) n, [](unsigned a, unsigned b) { return (a >> 16) < (b >> 16); });
  return 0;
}

}  // extern "C"
