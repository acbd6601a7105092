"""Thin Python mirror of the C-ABI context (include/c2g.h).  torch is plumbing only: device buffers, streams, NCCL.

`Engine` owns one c2g context on one GPU.  Names follow the reference: a *scan* is one ContourManager
(include/cont2/contour_mng.h:414), a *slot* is where its finished descriptor lives in HBM, gidx == slot for DB scans.
"""
import ctypes as C

import numpy as np

from . import capi
from . import ctypes_defs as D


class Engine:
    def __init__(self, cm_cfg: D.CmConfig = None, db_cfg: D.DbConfig = None, device: int = 0, scan_capacity: int = 1024,
                 max_batch: int = 256, max_points: int = 256 * 131072):
        self.cm_cfg = cm_cfg or D.kitti_cm_config()
        self.db_cfg = db_cfg or D.kitti_db_config()
        self.device = device
        self.scan_capacity = scan_capacity
        self.max_batch = max_batch
        self.n_cells = self.cm_cfg.n_row * self.cm_cfg.n_col
        h = C.c_void_p()
        capi.check(capi.lib().c2g_create(C.byref(self.cm_cfg), C.byref(self.db_cfg), device, scan_capacity, max_batch,
                                         max_points, C.byref(h)), "c2g_create")
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            capi.lib().c2g_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- streams ---------------------------------------------------------------------------------------------------
    def set_stream(self, cuda_stream_ptr: int):
        capi.check(capi.lib().c2g_set_stream(self.h, C.c_void_p(cuda_stream_ptr)), "c2g_set_stream")

    def sync(self):
        capi.check(capi.lib().c2g_sync(self.h), "c2g_sync")

    # ---- ingest (ContourManager::makeBEV + makeContoursRecurs) --------------------------------------------------------
    def ingest(self, pts, offsets, first_slot: int = 0, int_ids=None, on_device: bool = None):
        """pts: float32 [sum_n, 4] numpy array / torch tensor (host or cuda); offsets: int64 [B+1] in points."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        B = len(offsets) - 1
        if on_device is None:
            on_device = bool(getattr(pts, "is_cuda", False))
        ids = None if int_ids is None else np.ascontiguousarray(int_ids, np.int32)
        capi.check(capi.lib().c2g_ingest(self.h, capi.ptr(pts), capi.ptr(offsets), B, int(on_device), first_slot,
                                         capi.ptr(ids)), "c2g_ingest")
        return B

    def ingest_xyz(self, xyz, offsets, first_slot: int = 0, int_ids=None, on_device: bool = None):
        """Like ingest with 12 bytes per point: xyz is float32 [sum_n, 3] (the intensity column of the .bin record dropped)."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        B = len(offsets) - 1
        if on_device is None:
            on_device = bool(getattr(xyz, "is_cuda", False))
        ids = None if int_ids is None else np.ascontiguousarray(int_ids, np.int32)
        capi.check(capi.lib().c2g_ingest_xyz(self.h, capi.ptr(xyz), capi.ptr(offsets), B, int(on_device), first_slot,
                                             capi.ptr(ids)), "c2g_ingest_xyz")
        return B

    def ingest_bev_only(self, pts, offsets, on_device: bool = None):
        offsets = np.ascontiguousarray(offsets, np.int64)
        if on_device is None:
            on_device = bool(getattr(pts, "is_cuda", False))
        capi.check(capi.lib().c2g_ingest_bev_only(self.h, capi.ptr(pts), capi.ptr(offsets), len(offsets) - 1,
                                                  int(on_device)), "c2g_ingest_bev_only")

    def heads(self, first_slot: int, n: int) -> np.ndarray:
        out = np.zeros(n, D.SCAN_HEAD_DTYPE)
        capi.check(capi.lib().c2g_get_heads(self.h, first_slot, n, capi.ptr(out)), "c2g_get_heads")
        return out

    def views(self, slot: int, head=None):
        """Per-level sorted ContourView lists of one scan: list of NLEV structured arrays."""
        raw = np.zeros(D.VIEW_CAP, D.VIEW_DTYPE)
        capi.check(capi.lib().c2g_get_views(self.h, slot, capi.ptr(raw)), "c2g_get_views")
        if head is None:
            head = self.heads(slot, 1)[0]
        return [raw[head["view_off"][l]: head["view_off"][l] + head["n_views"][l]].copy() for l in range(D.NLEV)]

    def bev(self, batch_index: int):
        b, r, c = (np.empty(self.n_cells, np.float32) for _ in range(3))
        capi.check(capi.lib().c2g_get_bev(self.h, batch_index, capi.ptr(b), capi.ptr(r), capi.ptr(c)), "c2g_get_bev")
        return b, r, c

    def bev_compact(self, batch_index: int, full_tile_variant: bool = False):
        """Scatter -> contour kernel hand-off of one scan of the last batch: (planes [NLEV, n_row, ceil(n_col/32)] uint32,
        fg [n_fg, 4] float32 = (height, row_f, col_f, 0) in raster order, n_occupied)."""
        wpr = (self.cm_cfg.n_col + 31) // 32
        planes = np.zeros((D.NLEV, self.cm_cfg.n_row, wpr), np.uint32)
        fg = np.zeros((self.n_cells, 4), np.float32)
        hdr = np.zeros(2, np.int32)
        capi.check(capi.lib().c2g_get_bev_compact(self.h, batch_index, int(full_tile_variant), capi.ptr(planes), capi.ptr(fg),
                                                  capi.ptr(hdr)), "c2g_get_bev_compact")
        return planes, fg[: int(hdr[1])].copy(), int(hdr[0])

    def scatter_deferred(self) -> int:
        """Scans of the last scatter launch that the fast kernel handed to the general 64-bit kernel."""
        n = np.zeros(1, np.int32)
        capi.check(capi.lib().c2g_scatter_deferred(self.h, capi.ptr(n)), "c2g_scatter_deferred")
        return int(n[0])

    def tiles(self, batch_index: int):
        t = np.empty(self.n_cells, np.uint64)
        capi.check(capi.lib().c2g_get_tiles(self.h, batch_index, capi.ptr(t)), "c2g_get_tiles")
        return t

    def copy_slots(self, src_first: int, dst_first: int, n: int):
        capi.check(capi.lib().c2g_copy_slots(self.h, src_first, dst_first, n), "c2g_copy_slots")

    # ---- ContourDB::addScan / pushAndBalance (host-side LayerDB bookkeeping inside the library) ----------------------
    def db_add_scans(self, first_slot: int, n: int, ts):
        ts = np.ascontiguousarray(ts, np.float64)
        assert len(ts) == n
        capi.check(capi.lib().c2g_db_add_scans(self.h, first_slot, n, capi.ptr(ts)), "c2g_db_add_scans")

    def db_push_and_balance(self, seed: int, ts: float):
        capi.check(capi.lib().c2g_db_push_and_balance(self.h, seed, float(ts)), "c2g_db_push_and_balance")

    # ---- windowed online loop (query -> addScan -> pushAndBalance for W consecutive scans) ---------------------------------
    def online_stage(self, pts, offsets, int_ids=None, on_device: bool = None, xyz: bool = False):
        """Ingest the next window into slots db_size.. and start reading its keys back (asynchronous).  xyz: 12 B / point input."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        if on_device is None:
            on_device = bool(getattr(pts, "is_cuda", False))
        ids = None if int_ids is None else np.ascontiguousarray(int_ids, np.int32)
        fn = capi.lib().c2g_online_stage_xyz if xyz else capi.lib().c2g_online_stage
        capi.check(fn(self.h, capi.ptr(pts), capi.ptr(offsets), len(offsets) - 1, int(on_device), capi.ptr(ids)), "c2g_online_stage")

    def online_commit(self, ts, seeds, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble, results_out):
        """Bookkeeping + queries of the oldest staged window; results_out (numpy QUERY_RESULT_DTYPE array or pinned torch uint8
        tensor) is filled asynchronously: call sync() before reading it.  ts / seeds must stay alive only for the call."""
        ts = np.ascontiguousarray(ts, np.float64)
        seeds = np.ascontiguousarray(seeds, np.int32)
        capi.check(capi.lib().c2g_online_commit(self.h, capi.ptr(ts), capi.ptr(seeds), C.byref(lb), C.byref(ub),
                                                capi.ptr(results_out)), "c2g_online_commit")

    def online_window(self, pts, offsets, ts, seeds, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble, int_ids=None, on_device: bool = None):
        """W iterations of query -> addScan -> pushAndBalance, results identical to the scan-by-scan calls."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        W = len(offsets) - 1
        if on_device is None:
            on_device = bool(getattr(pts, "is_cuda", False))
        ids = None if int_ids is None else np.ascontiguousarray(int_ids, np.int32)
        ts = np.ascontiguousarray(ts, np.float64)
        seeds = np.ascontiguousarray(seeds, np.int32)
        assert len(ts) == W and len(seeds) == W
        res = np.zeros(W, D.QUERY_RESULT_DTYPE)
        capi.check(capi.lib().c2g_online_window(self.h, capi.ptr(pts), capi.ptr(offsets), W, int(on_device), capi.ptr(ids),
                                                capi.ptr(ts), capi.ptr(seeds), C.byref(lb), C.byref(ub), capi.ptr(res)),
                   "c2g_online_window")
        self.check_results(res)
        return res

    def online_host_seconds(self) -> dict:
        """Host seconds spent so far inside online_commit, by part (measurement aid)."""
        out = np.zeros(4, np.float64)
        capi.check(capi.lib().c2g_online_host_seconds(self.h, capi.ptr(out)), "c2g_online_host_seconds")
        return dict(zip(("layerdb_bookkeeping", "knn_launches", "mirror_patches", "chain_launches"), (float(v) for v in out)))

    def online_runs(self) -> int:
        return int(capi.lib().c2g_online_runs(self.h))

    def online_groups(self) -> int:
        return int(capi.lib().c2g_online_groups(self.h))

    def db_size(self) -> int:
        return int(capi.lib().c2g_db_size(self.h))

    def db_sync(self):
        capi.check(capi.lib().c2g_db_sync(self.h), "c2g_db_sync")

    def db_layer_state(self, ll: int):
        rng = np.zeros(D.NUM_BUCKETS + 1, np.float32)
        ts = np.zeros(D.NUM_BUCKETS, np.int32)
        bs = np.zeros(D.NUM_BUCKETS, np.int32)
        capi.check(capi.lib().c2g_db_layer_state(self.h, ll, capi.ptr(rng), capi.ptr(ts), capi.ptr(bs)), "c2g_db_layer_state")
        return rng, ts, bs

    def db_bucket_tree(self, ll: int, bucket: int):
        n = int(self.db_layer_state(ll)[1][bucket])
        keys = np.zeros((n, D.KEY_DIM), np.float32)
        gidx = np.zeros(n, np.int32)
        seq = np.zeros(n, np.int32)
        if n:
            capi.check(capi.lib().c2g_db_bucket_tree(self.h, ll, bucket, capi.ptr(keys), capi.ptr(gidx), capi.ptr(seq)),
                       "c2g_db_bucket_tree")
        return keys, gidx, seq

    # ---- database mirror + query -----------------------------------------------------------------------------------------
    def db_set_layer(self, ll: int, keys: np.ndarray, gidx: np.ndarray, seq: np.ndarray, bucket: np.ndarray,
                     bucket_ranges: np.ndarray):
        keys = np.ascontiguousarray(keys, np.float32).reshape(-1, D.KEY_DIM)
        gidx = np.ascontiguousarray(gidx, np.int32)
        seq = np.ascontiguousarray(seq, np.int8)
        bucket = np.ascontiguousarray(bucket, np.uint8)
        rng = np.ascontiguousarray(bucket_ranges, np.float32)
        assert rng.shape == (D.NUM_BUCKETS + 1,)
        capi.check(capi.lib().c2g_db_set_layer(self.h, ll, keys.shape[0], capi.ptr(keys), capi.ptr(gidx), capi.ptr(seq),
                                               capi.ptr(bucket), capi.ptr(rng)), "c2g_db_set_layer")

    def hint_slots(self, B: int) -> int:
        return B * self.db_cfg.n_q_levels * D.MAX_PIV * self.db_cfg.nnk

    def query(self, first_slot: int, B: int, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble, want_trace: bool = False):
        res = np.zeros(B, D.QUERY_RESULT_DTYPE)
        hints = scores = None
        if want_trace:
            hints = np.zeros(self.hint_slots(B), D.HINT_DTYPE)
            scores = np.zeros(self.hint_slots(B), D.PAIR_SCORE_DTYPE)
        capi.check(capi.lib().c2g_query(self.h, first_slot, B, C.byref(lb), C.byref(ub), capi.ptr(res), capi.ptr(hints),
                                        capi.ptr(scores)), "c2g_query")
        self.check_results(res)
        return (res, hints, scores) if want_trace else res

    @staticmethod
    def check_results(res: np.ndarray, allow_overflow: bool = False):
        """The reference's CandidateManager keeps every candidate pose (contour_db.h:352-363); the device keeps C2G_MAX_CAND per
        query scan and flags the rest.  A dropped pose can change which loop closure is returned, so an overflow is an error,
        like a refinement whose pair list did not fit (fine_flags)."""
        if not allow_overflow and int(res["overflow"].max(initial=0)) != 0:
            bad = np.nonzero(res["overflow"])[0]
            raise capi.C2gError(f"query scan(s) {bad[:8].tolist()} proposed more than {D.MAX_CAND} candidate poses (C2G_MAX_CAND): "
                                "results would differ from the reference")
        for r in res:
            n = min(int(r["n_cand"]), D.MAX_CAND)
            if n and int(r["cand"][:n]["fine_flags"].max()) != 0:
                raise capi.C2gError("a refinement's pre-selected pair list overflowed the device scratch")

    def query_async(self, first_slot: int, B: int, lb: D.ScoreEnsemble, ub: D.ScoreEnsemble):
        capi.check(capi.lib().c2g_query_async(self.h, first_slot, B, C.byref(lb), C.byref(ub)), "c2g_query_async")

    def query_buffers(self):
        r, h, s, n = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_longlong()
        capi.check(capi.lib().c2g_query_buffers(self.h, C.byref(r), C.byref(h), C.byref(s), C.byref(n)), "c2g_query_buffers")
        return r.value, h.value, s.value, n.value

    def query_export(self, B: int, hints_dst=None, scores_dst=None, results_dst_host=None):
        """Async copies of the last query's buffers: device tensors for hints / scores, a pinned host tensor for results."""
        capi.check(capi.lib().c2g_query_export(self.h, B, capi.ptr(hints_dst), capi.ptr(scores_dst), capi.ptr(results_dst_host)),
                   "c2g_query_export")

    def finish_from_scores(self, first_slot: int, B: int, lb: D.ScoreEnsemble, hints_dev: int, scores_dev: int):
        res = np.zeros(B, D.QUERY_RESULT_DTYPE)
        capi.check(capi.lib().c2g_finish_from_scores(self.h, first_slot, B, C.byref(lb), C.c_void_p(hints_dev),
                                                     C.c_void_p(scores_dev), capi.ptr(res)), "c2g_finish_from_scores")
        return res

    QUERY_KERNELS = ("knn", "prefilter", "score", "replay", "gmm_gate", "output", "refine", "rank")

    def query_profile(self, enable: bool = True, read: bool = False):
        """Per-kernel CUDA-event timing of query_async (ms per kernel of the last profiled call when read=True)."""
        ms = np.zeros(8, np.float32) if read else None
        capi.check(capi.lib().c2g_query_profile(self.h, int(enable), capi.ptr(ms)), "c2g_query_profile")
        return dict(zip(self.QUERY_KERNELS, (float(v) for v in ms))) if read else None

    WORK_COUNTERS = ("knn_keys_evaluated", "knn_boxes_tested", "gate_preselect_tests", "gate_terms", "refine_preselect_tests",
                     "refine_terms", "refine_evaluations", "spare")

    def work_counters(self, enable: bool = True) -> dict:
        """Read + clear the query kernels' work counters, then enable / disable counting (c2g_work_counters)."""
        out = np.zeros(8, np.uint64)
        capi.check(capi.lib().c2g_work_counters(self.h, int(enable), capi.ptr(out)), "c2g_work_counters")
        return dict(zip(self.WORK_COUNTERS, (int(v) for v in out)))

    def exp_mode(self) -> int:
        """Which glibc exp() variant the device reproduces (0 = none matched the host libm: libdevice exp, keys may differ by 1 ulp)."""
        return int(capi.lib().c2g_exp_mode(self.h))

    def launch_count(self) -> int:
        return int(capi.lib().c2g_launch_count(self.h))
