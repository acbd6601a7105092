"""CPU-only, world_size 2 over gloo: the N > 1 host logic of the query path — query sharding, the ONE all-gather of the
per-query result records (c2g_query_result) and the foreign-block verification bench.py runs on hardware.  The records come
from the oracle (each rank queries its own shard of the query scans against its replica), so this also checks that sharding
by query leaves every per-query result unchanged (DYNAMIC_THRES=0 makes hint checks independent, CMakeLists.txt:21)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_q, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from contour_context_b200 import ctypes_defs as D
    from contour_context_b200 import multi, synth
    from helpers import make_batch
    from oracle import c2o

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg, dbc = D.kitti_cm_config(), D.kitti_db_config()
    lb, ub = D.kitti_thres()
    n_pts = 40000
    # every rank builds the same (replicated) 12-scan DB
    seeds, visits = synth.db_layout(12, 3, first_scene=500)
    pts, off = make_batch(seeds, visits, n_pts)
    db = c2o.DB(dbc)
    for i in range(12):
        db.add_scan(c2o.Scan(cfg, i).ingest(pts[off[i]:off[i + 1]]), 0.1 * i)
    for k in range(12):
        db.push_and_balance(k, 1000.0 + k)
    qpts, qoff = make_batch([500 + i % 4 for i in range(n_q)], [3 + i // 4 for i in range(n_q)], n_pts, noise_seed=5)

    def record(qi):
        res, _, _ = db.query(c2o.Scan(cfg, 100 + qi).ingest(qpts[qoff[qi]:qoff[qi + 1]]), lb, ub)
        return res

    beg, end = multi.shard_range(n_q, world, rank)
    assert end - beg == n_q // world
    loc = np.array([record(q) for q in range(beg, end)], D.QUERY_RESULT_DTYPE)
    r_loc = torch.from_numpy(loc.view(np.uint8).reshape(-1).copy())
    r_all = multi.split_gathered(multi.all_gather_records(r_loc, world), world, D.QUERY_RESULT_DTYPE)
    # every rank now holds the outcome of the whole batch: (1) it equals a single-rank run of ALL queries, (2) the block of the
    # other rank equals those queries recomputed on this rank's replica (the check bench.py runs under torch.distributed.run)
    full = np.array([record(q) for q in range(n_q)], D.QUERY_RESULT_DTYPE)
    ok = r_all.reshape(-1).tobytes() == full.tobytes()
    other = (rank + 1) % world
    ob, oe = multi.shard_range(n_q, world, other)
    ok = ok and multi.verify_foreign_block(r_all, other, full[ob:oe]) == 0
    tampered = r_all.copy()
    tampered[other][0]["n_cand"] += 1
    ok = ok and multi.verify_foreign_block(tampered, other, full[ob:oe]) == 1
    n_pass = len(multi.loop_closures(r_all.reshape(-1)))
    with open(os.path.join(out_dir, f"rank{rank}.txt"), "w") as f:
        f.write(f"{int(ok)} {n_pass}\n")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from contour_context_b200 import multi

    for n in (0, 1, 7, 8, 592, 1000):
        for w in (1, 2, 3, 8):
            cuts = [multi.shard_range(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gather_of_results_and_foreign_block_check(oracle, tmp_path):
    import torch.multiprocessing as mp

    port = _free_port()
    mp.spawn(_worker, args=(2, port, 4, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        ok, n_pass = open(tmp_path / f"rank{r}.txt").read().split()
        assert ok == "1"
        assert int(n_pass) > 0
