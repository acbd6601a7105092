#!/usr/bin/env python
"""bench.py — scans/sec of the cont2contops hot path (ingest + query against a 5 000-scan database), BASELINE.json's metric.

Workload (BASELINE.json configs[2], the configuration the metric string is quoted on): synthetic 120 000-point scans in
the KITTI .bin layout; a database of 5 000 scans (1 250 scenes x 4 visits) whose descriptors and retrieval-key tables are
resident in HBM (replicated on every rank: 5 000 scans are ~0.2 GB of descriptors); one *step* = one batch of Q query
scans per rank going through the whole path: BEV scatter -> contours/keys/BCI -> ranged kNN -> hint scoring cascade ->
proposal merge + GMM-L2 -> (N > 1) one NCCL all-gather of the per-pair score records.

  value : Q*N / step time with the query points already resident in HBM (CUDA events, max over ranks)
  e2e   : the same step through the public C-ABI call with HOST buffers: pinned host -> device copy of the points and
          device -> host read of the results inside the timed region
  --impl reference : the CPU restatement of the reference path (oracle/, linked against the reference's own nanoflann when
          oracle/_ref was built) on all host cores, same workload, bounded sample per step.

Weak scaling: per-rank work is fixed (Q queries per rank against the same 5 000-scan DB).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec ingest+query vs 5k-scan DB"
N_PTS = 120000
VISITS = 4


# DRAM traffic per 120k-point scan measured by `ncu --set full` (profiles/r1e_ncu_full_summary.csv, one 592-scan launch each):
# (dram__bytes_read.sum + dram__bytes_write.sum) / 592.  Algorithmic bytes are 1.92 MB (K1) and 1.98 MB (K1+K2) per scan.
NCU_DRAM_BYTES_PER_SCAN = {"bev_scatter_kernel": (1.136680e9 + 51.249920e6) / 592, "contour_kernel": (0.305138e9 + 87.707648e6) / 592}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--db-scans", type=int, default=5000)
    ap.add_argument("--queries", type=int, default=1184,
                    help="query scans per rank per step: SURVEY.md 8d config 3 asks for Q = 1 024; 8 x 148 SMs = 1 184 is the next multiple of the SM count")
    ap.add_argument("--points", type=int, default=N_PTS)
    ap.add_argument("--cpu-sample", type=int, default=96, help="query scans of the single-thread CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu_index = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for k, nm in enumerate(names):
                    if r[5 + k].lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
def build_db(eng, synth, torch, n_db, n_pts, chunk=148):
    """Untimed setup: ingest the DB scans on the GPU and run ContourDB::addScan / pushAndBalance for each (growing DB
    bookkeeping, ts_i = 0.1 i), then flush every buffered key into its tree."""
    t0 = time.time()
    for i0 in range(0, n_db, chunk):
        n = min(chunk, n_db - i0)
        seeds, visits = synth.db_layout(n_db, VISITS)
        pts = synth.make_scans(seeds[i0:i0 + n], visits[i0:i0 + n], n_pts, device="cuda", noise_seed=i0).reshape(-1, 4)
        offsets = np.arange(n + 1, dtype=np.int64) * n_pts
        torch.cuda.synchronize()  # the generator ran on torch's stream, the context has its own
        eng.ingest(pts, offsets, first_slot=i0, int_ids=np.arange(i0, i0 + n))
        eng.sync()
        ts = 0.1 * np.arange(i0, i0 + n)
        for j in range(n):
            eng.db_add_scans(i0 + j, 1, ts[j:j + 1])
            eng.db_push_and_balance(i0 + j, ts[j])
        del pts
    t_end = 0.1 * n_db + 525.0
    for k in range(16):
        eng.db_push_and_balance(k, t_end + k)
    eng.db_sync()
    return time.time() - t0


def make_queries(synth, torch, n_q, n_pts, first_scene, chunk=148):
    """Query scans = a NEW visit (index VISITS) of DB scenes, so every query has true loop-closure partners in the DB."""
    outs = []
    for i0 in range(0, n_q, chunk):
        n = min(chunk, n_q - i0)
        seeds = [first_scene + i0 + k for k in range(n)]
        outs.append(synth.make_scans(seeds, [VISITS] * n, n_pts, device="cuda", noise_seed=777 + i0))
    return torch.cat(outs).reshape(-1, 4).contiguous()


def run_b200(args):
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g
    from contour_context_b200 import capi
    from contour_context_b200 import ctypes_defs as D
    from contour_context_b200 import synth
    from contour_context_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner to stdout otherwise (the JSON line must be alone)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        g.build_c2g()
    if world > 1:
        dist.barrier()
    capi.lib()

    n_db, Q, n_pts = args.db_scans, args.queries, args.points
    n_scenes = n_db // VISITS
    eng = Engine(device=local_rank, scan_capacity=n_db + Q + 8, max_batch=max(Q, 148), max_points=Q * n_pts)
    stream = torch.cuda.Stream()
    eng.set_stream(stream.cuda_stream)
    lb, ub = D.kitti_thres()

    t_setup = build_db(eng, synth, torch, n_db, n_pts)
    # distinct query scenes per rank
    q_dev = make_queries(synth, torch, Q, n_pts, first_scene=(rank * Q) % max(1, n_scenes - Q))
    q_host = torch.empty(q_dev.shape, dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    torch.cuda.synchronize()
    offsets = np.arange(Q + 1, dtype=np.int64) * n_pts
    q_first = n_db
    res_host = np.zeros(Q, D.QUERY_RESULT_DTYPE)
    res_pinned = torch.empty(res_host.nbytes, dtype=torch.uint8, pin_memory=True)
    per_rank_hints = eng.hint_slots(Q)
    if world > 1:  # send / receive buffers of the all-gather (torch tensors: NCCL plumbing)
        sc_local = torch.empty(per_rank_hints * D.PAIR_SCORE_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        hi_local = torch.empty(per_rank_hints * D.HINT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        sc_all = torch.empty(world * sc_local.numel(), dtype=torch.uint8, device="cuda")
        hi_all = torch.empty(world * hi_local.numel(), dtype=torch.uint8, device="cuda")

    pending = []  # NCCL work handles of the previous step's all-gather

    def drain():
        with torch.cuda.stream(stream):
            while pending:
                pending.pop().wait()  # stream-level wait: `stream` continues after the collective has finished

    def step(host_inputs: bool):
        with torch.cuda.stream(stream):
            if host_inputs:
                eng.ingest(q_host, offsets, first_slot=q_first, on_device=False)   # H2D inside c2g_ingest
            else:
                eng.ingest(q_dev, offsets, first_slot=q_first, on_device=True)
            eng.query_async(q_first, Q, lb, ub)
            if world > 1:
                # the path's one exchange step: publish the per-pair score records to every rank.  The collective runs on
                # NCCL's stream and overlaps the NEXT step's ingest + kNN; the send buffers are reused only after it is done.
                drain()
                eng.query_export(Q, hi_local, sc_local, None)
                pending.append(dist.all_gather_into_tensor(sc_all, sc_local, async_op=True))
                pending.append(dist.all_gather_into_tensor(hi_all, hi_local, async_op=True))
            if host_inputs:
                eng.query_export(Q, None, None, res_pinned)                        # D2H of the step's results

    def timed(host_inputs: bool, steps: int, warmup: int):
        for _ in range(warmup):
            step(host_inputs)
        drain()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = eng.launch_count()
        e0.record(stream)
        for _ in range(steps):
            step(host_inputs)
        drain()  # the last step's all-gather is inside the timed region
        e1.record(stream)
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = eng.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(False, args.steps, args.warmup)
    ms_e2e, _ = timed(True, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None

    # per-kernel timings on rank 0 (CUDA events on the launching stream), for the roofline entry
    kern = {}
    if rank == 0:
        def ev_time(fn, reps=5):
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    a.record(stream)
                    fn()
                    b.record(stream)
                stream.synchronize()
                ts.append(a.elapsed_time(b))
            return float(np.mean(ts[1:])) if len(ts) > 1 else float(ts[0])

        kern["bev_scatter_ms"] = ev_time(lambda: eng.ingest_bev_only(q_dev, offsets, on_device=True))
        kern["ingest_ms"] = ev_time(lambda: eng.ingest(q_dev, offsets, first_slot=q_first, on_device=True))
        kern["contours_ms"] = kern["ingest_ms"] - kern["bev_scatter_ms"]
        kern["query_ms"] = ev_time(lambda: eng.query_async(q_first, Q, lb, ub))
        eng.query_profile(True)
        acc = {}
        for _ in range(3):
            with torch.cuda.stream(stream):
                eng.query_async(q_first, Q, lb, ub)
            for k, v in eng.query_profile(True, read=True).items():
                acc.setdefault(k, []).append(v)
        eng.query_profile(False)
        kern["query_kernels_ms"] = {k: float(np.mean(v[1:])) for k, v in acc.items()}

    # sanity: the timed work produced real loop closures (not measured; guards against timing an empty path)
    res = eng.query(q_first, Q, lb, ub)
    n_found = int((res["n_cand"] > 0).sum())
    heads_q = eng.heads(q_first, min(Q, 8))
    assert int(heads_q["status"].max()) == 0

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s"
        total_q = Q * world
        value = total_q / (ms_dev / args.steps) * 1e3
        e2e_val = total_q / (ms_e2e / args.steps) * 1e3
        # dominant kernels = the ingest pair (bev_scatter + contours): SURVEY.md §8d algorithmic bytes 1.98 MB per scan
        # (16 B x 120 000 points read + descriptor written); the BEV kernel alone moves 1.92 MB per scan.
        alg_bytes_ingest = Q * (16.0 * n_pts + 60000.0)
        ach = alg_bytes_ingest / (kern["ingest_ms"] * 1e-3) / 1e9
        ach_bev = Q * 16.0 * n_pts / (kern["bev_scatter_ms"] * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (BEV/CCL/keys, no FMA) + f64 (moments, GMM-L2)", "data": "synthetic",
            "config": {"workload": "configs[2]: synthetic 120k-pt scans, 5k-scan DB, batched ingest+query on 1 GPU per rank",
                       "db_scans": n_db, "queries_per_rank_per_step": Q, "points_per_scan": n_pts,
                       "parallelism": f"query batches sharded over {world} rank(s), DB replicated, 1 NCCL all-gather of pair scores",
                       "l2": "inputs larger than L2 (each step streams %.2f GB of points)" % (Q * n_pts * 16 / 1e9),
                       "refine": "fineOptimize's L-BFGS refinement of <=10 candidates per query included (refine.cu)"},
            "e2e": {"value": e2e_val, "unit": "scans/s", "h2d_bytes_per_step": int(Q * n_pts * 16 + (Q + 1) * 8),
                    "d2h_bytes_per_step": int(res_host.nbytes), "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                         "traffic": Q * (NCU_DRAM_BYTES_PER_SCAN["bev_scatter_kernel"] + NCU_DRAM_BYTES_PER_SCAN["contour_kernel"])
                         if n_pts == 120000 else None,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per scan from the ncu --set full capture in "
                                           "profiles/r1e_ncu_full_summary.csv (592-scan launch), scaled to this launch's scan count",
                         "kernel": "bev_scatter_kernel + contour_kernel (ingest pair, 1.98 MB algorithmic bytes per scan)",
                         "peak_source": peak_src,
                         "bev_scatter_only": {"achieved": ach_bev, "frac": ach_bev / peak, "ms": kern["bev_scatter_ms"]},
                         "kernel_ms": kern},
            "sanity": {"queries_with_loop_candidate": n_found, "of": int(Q), "db_build_s": t_setup, "exp_mode": eng.exp_mode()},
        }
    if world > 1:
        dist.barrier()
    eng_keep = eng  # keep the context alive until the CPU baseline has read the descriptors
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args, eng_keep, q_host, offsets, lb, ub)
    if rank == 0:
        print(json.dumps(out))
    eng.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------------------------
def cpu_baseline(args, eng, q_host, offsets, lb, ub):
    """Single-thread CPU port (oracle, + the reference's own nanoflann KD-tree when oracle/_ref is present) on a bounded
    sample of the same workload: the DB is rebuilt on the CPU side from the descriptors the GPU path produced (descriptor
    parity is established by tests/), then `cpu_sample` query scans go through ingest + query."""
    from contour_context_b200 import ctypes_defs as D
    from oracle import c2o

    nf = c2o.use_nanoflann_build()
    n_db = args.db_scans
    odb = c2o.DB(eng.db_cfg)
    t0 = time.time()
    B = 128
    for i0 in range(0, n_db, B):
        n = min(B, n_db - i0)
        heads = eng.heads(i0, n)
        for j in range(n):
            raw = np.zeros(D.VIEW_CAP, D.VIEW_DTYPE)
            from contour_context_b200 import capi
            capi.check(capi.lib().c2g_get_views(eng.h, i0 + j, capi.ptr(raw)))
            s = c2o.Scan.from_descriptor(eng.cm_cfg, heads[j], raw)
            odb.add_scan(s, 0.1 * (i0 + j))
            odb.push_and_balance(i0 + j, 0.1 * (i0 + j))
    t_end = 0.1 * n_db + 525.0
    for k in range(16):
        odb.push_and_balance(k, t_end + k)
    t_build = time.time() - t0
    S = min(args.cpu_sample, len(offsets) - 1)
    pts = q_host.numpy()[: offsets[S]]
    t1 = time.time()
    res, stages = c2o.run_loop(odb, eng.cm_cfg, pts, offsets[: S + 1], 100000, np.zeros(S), True, False, lb, ub)
    dt = time.time() - t1
    return {"value": S / dt, "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": f"{S} query scans of the same batch, ingest+query vs the same {n_db}-scan DB (DB rebuilt from the GPU "
                      f"descriptors in {t_build:.1f} s, untimed); kNN = " + ("reference's vendored nanoflann" if nf else "exhaustive scan"),
            "stage_ms_per_scan": {"make bev": stages[0] / S * 1e3, "KNN search": stages[1] / S * 1e3,
                                  "Constell": stages[2] / S * 1e3, "L2 opt": stages[3] / S * 1e3},
            "host_cores_available": os.cpu_count()}


def run_reference(args):
    """--impl reference: the reference's CPU path (restated in oracle/, KD-tree = the reference's vendored nanoflann when
    oracle/_ref exists) on all host cores. Rank 0 only; other ranks exit."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    from contour_context_b200 import ctypes_defs as D
    from contour_context_b200 import synth
    from oracle import c2o

    nf = c2o.use_nanoflann_build()
    cfg, dbc = D.kitti_cm_config(), D.kitti_db_config()
    lb, ub = D.kitti_thres()
    n_db, n_pts = args.db_scans, args.points
    threads = os.cpu_count() or 1
    dev = "cuda" if torch.cuda.is_available() else "cpu"  # torch only generates the synthetic input data here
    odb = c2o.DB(dbc)
    from concurrent.futures import ThreadPoolExecutor

    pool = ThreadPoolExecutor(threads)
    t0 = time.time()
    seeds, visits = synth.db_layout(n_db, VISITS)
    chunk = 148
    for i0 in range(0, n_db, chunk):
        n = min(chunk, n_db - i0)
        pts = synth.make_scans(seeds[i0:i0 + n], visits[i0:i0 + n], n_pts, device=dev, noise_seed=i0).cpu().numpy()
        scans = list(pool.map(lambda j: c2o.Scan(cfg, i0 + j).ingest(pts[j]), range(n)))
        for j, s in enumerate(scans):
            odb.add_scan(s, 0.1 * (i0 + j))
            odb.push_and_balance(i0 + j, 0.1 * (i0 + j))
    for k in range(16):
        odb.push_and_balance(k, 0.1 * n_db + 525.0 + k)
    t_build = time.time() - t0
    S = max(threads, min(args.queries, 4 * threads))  # bounded sample per step
    n_scenes = n_db // VISITS
    q = synth.make_scans(list(range(S)), [VISITS] * S, n_pts, device=dev, noise_seed=777).cpu().numpy().reshape(-1, 4)
    offsets = np.arange(S + 1, dtype=np.int64) * n_pts
    parts = np.array_split(np.arange(S), threads)

    def work(idx):
        if len(idx) == 0:
            return
        a, b = idx[0], idx[-1] + 1
        c2o.run_loop(odb, cfg, q[offsets[a]:offsets[b]], offsets[a:b + 1] - offsets[a], 100000 + a, np.zeros(b - a), True, False, lb, ub)

    def step():
        list(pool.map(work, parts))

    for _ in range(args.warmup):
        step()
    t1 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t1
    value = S * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32 + f64 (CPU, no FMA contraction)", "data": "synthetic",
        "config": {"workload": "configs[2]: synthetic 120k-pt scans, 5k-scan DB", "db_scans": n_db, "points_per_scan": n_pts,
                   "queries_per_step": S},
        "cpu_baseline": {"value": value, "unit": "scans/s", "cores": threads, "kind": "port",
                         "sample": f"{S} query scans per step over {threads} threads vs the {n_db}-scan DB (built on the CPU in {t_build:.0f} s); "
                                   + ("kNN through the reference's own vendored nanoflann (oracle/_ref)" if nf else "exhaustive-scan kNN")},
        "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
