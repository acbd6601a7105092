"""Full-size (BASELINE.json configs[2]: 120k-point scans, 5 000-scan DB) checks through size-independent properties — the CPU
oracle would need minutes for this DB, so nothing here compares against it:
  * determinism: the same batch ingested and queried twice gives byte-identical descriptors, hints, scores and results;
  * batch independence: a query scan processed alone (B = 1, another slot) gets the same hints / scores / result as in
    the 148-scan batch;
  * kNN invariants: per query key the kept squared distances ascend, stay below dist_ub and never repeat a (scan, contour);
  * self-retrieval: an exact copy of a DB scan retrieves that scan with refined correlation ~ 1 and identity transform;
  * revisits: a new visit of a DB scene retrieves a scan of that scene."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth

pytestmark = pytest.mark.gpu

N_DB, N_PTS, VISITS, NQ = 5000, 120000, 4, 148


@pytest.fixture(scope="module")
def big(built_lib):
    import torch

    from contour_context_b200.engine import Engine

    eng = Engine(scan_capacity=N_DB + 2 * NQ + 8, max_batch=NQ, max_points=NQ * N_PTS)
    seeds, visits = synth.db_layout(N_DB, VISITS)
    keep = {}
    for i0 in range(0, N_DB, NQ):
        n = min(NQ, N_DB - i0)
        pts = synth.make_scans(seeds[i0:i0 + n], visits[i0:i0 + n], N_PTS, device="cuda", noise_seed=i0).reshape(-1, 4)
        torch.cuda.synchronize()
        offsets = np.arange(n + 1, dtype=np.int64) * N_PTS
        eng.ingest(pts, offsets, first_slot=i0, int_ids=np.arange(i0, i0 + n), on_device=True)
        eng.sync()
        if i0 == 0:
            keep["copy_pts"] = pts[:8 * N_PTS].clone()  # DB scans 0..7, re-used as "exact copy" queries
        for j in range(n):
            t = np.array([0.1 * (i0 + j)])
            eng.db_add_scans(i0 + j, 1, t)
            eng.db_push_and_balance(i0 + j, float(t[0]))
    for k in range(16):
        eng.db_push_and_balance(k, 0.1 * N_DB + 525.0 + k)
    eng.db_sync()
    n_scenes = N_DB // VISITS
    q_scenes = [int(s) for s in np.linspace(0, n_scenes - 1, NQ - 8).astype(int)]
    q = synth.make_scans(q_scenes, [VISITS] * len(q_scenes), N_PTS, device="cuda", noise_seed=777).reshape(-1, 4)
    q = torch.cat([keep["copy_pts"], q])
    torch.cuda.synchronize()
    yield dict(eng=eng, q=q, q_scenes=q_scenes, torch=torch)
    eng.close()


def test_fullsize_properties(big):
    eng, q, torch = big["eng"], big["q"], big["torch"]
    lb, ub = D.kitti_thres()
    offsets = np.arange(NQ + 1, dtype=np.int64) * N_PTS
    first = N_DB
    eng.ingest(q, offsets, first_slot=first, on_device=True)
    heads_a, res_a, hints_a, scores_a = eng.heads(first, NQ).copy(), *eng.query(first, NQ, lb, ub, want_trace=True)
    # --- determinism
    eng.ingest(q, offsets, first_slot=first, on_device=True)
    heads_b, (res_b, hints_b, scores_b) = eng.heads(first, NQ), eng.query(first, NQ, lb, ub, want_trace=True)
    assert heads_a.tobytes() == heads_b.tobytes()
    assert hints_a.tobytes() == hints_b.tobytes() and scores_a.tobytes() == scores_b.tobytes() and res_a.tobytes() == res_b.tobytes()
    assert int(heads_a["status"].max()) == 0 and int(res_a["overflow"].max()) == 0
    # --- batch independence: scans 3, 40, 147 alone, in a different slot
    per_q = eng.hint_slots(1)
    alone = first + NQ
    for j in (3, 40, 147):
        eng.ingest(q[j * N_PTS:(j + 1) * N_PTS], np.array([0, N_PTS], np.int64), first_slot=alone, on_device=True)
        r1, h1, s1 = eng.query(alone, 1, lb, ub, want_trace=True)
        hj, sj = hints_a[j * per_q:(j + 1) * per_q].copy(), scores_a[j * per_q:(j + 1) * per_q]
        hj["q_idx"] = 0
        assert h1.tobytes() == hj.tobytes() and s1.tobytes() == sj.tobytes(), j
        assert r1[0].tobytes() == res_a[j].tobytes(), j
    # --- kNN invariants
    nnk = eng.db_cfg.nnk
    hh = hints_a.reshape(NQ, eng.db_cfg.n_q_levels, D.MAX_PIV, nnk)
    valid = hh["cand_gidx"] >= 0
    d = np.where(valid, hh["dist_sq"], np.float32(3e38))
    assert (np.diff(d, axis=-1) >= 0).all()  # ascending; invalid slots trail
    assert (hh["cand_gidx"][valid] < N_DB).all()
    ident = hh["cand_gidx"].astype(np.int64) * 16 + hh["cand_seq"]
    for qi in range(0, NQ, 13):
        for ll in range(eng.db_cfg.n_q_levels):
            for s in range(D.MAX_PIV):
                v = ident[qi, ll, s][valid[qi, ll, s]]
                assert len(np.unique(v)) == len(v)
    keys = heads_a["keys"]
    for qi in range(0, NQ, 29):
        for ll in range(eng.db_cfg.n_q_levels):
            lev = eng.db_cfg.q_levels[ll]
            for s in range(D.MAX_PIV):
                k = keys[qi][lev][s].astype(np.float64)
                ub_d = (0.25 * k[0]) ** 2 + (0.25 * k[1]) ** 2 + (k[2] * (1 / 0.6 - 1)) ** 2
                dv = hh["dist_sq"][qi, ll, s][valid[qi, ll, s]]
                assert (dv <= ub_d * (1 + 1e-5)).all()
    # --- self-retrieval of exact copies of DB scans 0..7
    for j in range(8):
        r = res_a[j]
        assert r["n_cand"] >= 1
        c = r["cand"][0]
        assert int(c["cand_gidx"]) == j, (j, int(c["cand_gidx"]))
        assert c["corr_fine"] > 0.95 and abs(c["T_fine"][0] - 1) < 1e-4 and abs(c["T_fine"][1]) < 1e-3
        assert abs(c["T_fine"][2]) < 0.05 and abs(c["T_fine"][3]) < 0.05
    # --- revisits retrieve their own scene
    hits = 0
    for j, sc in enumerate(big["q_scenes"]):
        r = res_a[8 + j]
        if r["n_cand"] >= 1 and int(r["cand"][0]["cand_gidx"]) // VISITS == sc:
            hits += 1
    assert hits >= int(0.9 * len(big["q_scenes"])), hits
