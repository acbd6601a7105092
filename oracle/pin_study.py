"""ORACLE — TEST INFRASTRUCTURE ONLY.  Numbers behind DESIGN.md §2 "residual risk of the unpinned pieces".

The reference cannot be built here (Eigen / Ceres absent), so two restated third-party pieces have no reference-held pin:
 (i)  Eigen::SelfAdjointEigenSolver<Matrix2f> (include/cont2/contour.h:165) -> eig_vals_ feed key[0], key[1]
      (include/cont2/contour_mng.h:813-816) bit for bit;
 (ii) Ceres' L-BFGS + Wolfe line search (include/cont2/correlation.h:206-238) -> the final correlation / SE(2).
This module bounds the risk with numbers:
 (i)  the restated solver against an exact (80-bit long double, then rounded once to float) eigen-decomposition over >= 10^6
      covariance matrices drawn from real descriptors: ulp error of each eigenvalue and the fraction of matrices for which the
      key entries sqrt(lambda * cnt) differ in any bit from the correctly rounded ones;
 (ii) the restated solver against scipy's L-BFGS-B and BFGS on real candidate problems (revisits of synthetic scenes, start
      = the constellation transform of the candidate): distribution of |delta correlation|, and how often a decision at the
      shipped correlation threshold would change.

    python -m oracle.pin_study            # prints both tables (about a minute)
tests/test_oracle_pin_study.py runs reduced versions and asserts the bounds quoted in DESIGN.md.
"""
import ctypes as C

import numpy as np

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth
from oracle import c2o


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ulp_f32(a, b):
    a = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)  # positive finite floats only


def real_covariances(n_scans=24, n_pts=60000, seed0=40):
    """(c00, c01, c11, cnt) of every contour view with cnt >= 4 of n_scans synthetic scans (the oracle's own descriptors)."""
    cfg = D.kitti_cm_config()
    out = []
    for s0 in range(0, n_scans, 4):
        pts = synth.make_scans([seed0 + s0 + k for k in range(4)], [k % 3 for k in range(4)], n_pts, "cpu", seed0 + s0).numpy()
        for k in range(4):
            sc = c2o.Scan(cfg, s0 + k).ingest(np.ascontiguousarray(pts[k]))
            for lev in range(D.NLEV):
                v = sc.views(lev)
                v = v[v["cell_cnt"] >= 4]
                if len(v):
                    out.append(np.stack([v["pos_cov"][:, 0], v["pos_cov"][:, 1], v["pos_cov"][:, 3], v["cell_cnt"].astype(np.float32)], 1))
    return np.concatenate(out).astype(np.float32)


def eig_study(n_total=1_000_000, n_scans=24, rng_seed=3):
    """Restated Eigen 2x2 solver vs exact eigenvalues on n_total covariance matrices: the real ones plus jittered copies
    (every entry scaled by 1 + 2e-3 u, u ~ U(-1, 1), which keeps them in the population of contour covariances)."""
    L = c2o.lib()
    base = real_covariances(n_scans)
    rng = np.random.default_rng(rng_seed)
    reps = int(np.ceil(n_total / len(base)))
    M = np.tile(base, (reps, 1))[:n_total].copy()
    jit = 1.0 + 2e-3 * rng.uniform(-1, 1, (n_total, 3))
    jit[: len(base)] = 1.0  # the real matrices themselves come first, untouched
    M[:, :3] = (M[:, :3].astype(np.float64) * jit).astype(np.float32)
    # positive semi-definiteness can be lost by the jitter of near-singular matrices: keep those, Eigen does not care either
    ev = np.zeros((n_total, 2), np.float32)
    tmp_ev = np.zeros(2, np.float32)
    tmp_vec = np.zeros(4, np.float32)
    f = L.c2o_eig2f
    pe, pv = _p(tmp_ev), _p(tmp_vec)
    for i in range(n_total):
        f(float(M[i, 0]), float(M[i, 1]), float(M[i, 2]), pe, pv)
        ev[i] = tmp_ev
    a, b, c = (M[:, k].astype(np.longdouble) for k in range(3))
    half, d = (a + c) / 2, (a - c) / 2
    r = np.sqrt(d * d + b * b)
    exact = np.stack([half - r, half + r], 1)
    exact_f = exact.astype(np.float32)
    ok = (exact_f > 0).all(1) & np.isfinite(ev).all(1) & (ev > 0).all(1)
    ulp = _ulp_f32(ev[ok], exact_f[ok])
    unclamped = exact_f[ok][:, 0] >= np.float32(1.0)  # the smaller eigenvalue only matters when it survives the clamp below
    # what reaches the keys: eigenvalues clamped to >= point_sigma (contour.h:167-170), key = sqrt(lambda * cnt) in float
    ps = np.float32(1.0)
    cnt = M[ok, 3]
    k_ours = np.sqrt(np.maximum(ev[ok], ps) * cnt[:, None]).astype(np.float32)
    k_exact = np.sqrt(np.maximum(exact_f[ok], ps) * cnt[:, None]).astype(np.float32)
    key_ulp = _ulp_f32(k_ours, k_exact)
    return {
        "matrices": int(n_total), "real_matrices": int(len(base)), "compared": int(ok.sum()),
        "eig_ulp_max": {"lambda0_all": int(ulp[:, 0].max()), "lambda0_unclamped": int(ulp[unclamped, 0].max()) if unclamped.any() else 0,
                        "lambda1": int(ulp[:, 1].max())},
        "eig_ulp_mean": {"lambda0_unclamped": float(ulp[unclamped, 0].mean()) if unclamped.any() else 0.0, "lambda1": float(ulp[:, 1].mean())},
        "eig_exact_frac": {"lambda0_unclamped": float((ulp[unclamped, 0] == 0).mean()) if unclamped.any() else 1.0,
                           "lambda1": float((ulp[:, 1] == 0).mean())},
        "eig_rel_err_max": float((np.abs(ev[ok].astype(np.float64) - exact[ok].astype(np.float64)) / exact[ok].astype(np.float64).max(1, keepdims=True)).max()),
        "key_bits_differ_frac": [float((key_ulp[:, 1] > 0).mean()), float((key_ulp[:, 0] > 0).mean())],  # key[0] uses the larger eigenvalue
        "key_ulp_max": [int(key_ulp[:, 1].max()), int(key_ulp[:, 0].max())],
        "clamped_frac": float((exact_f[ok][:, 0] < ps).mean()),
    }


def refine_problems(n_scenes=60, n_pts=60000, first_scene=900):
    """Real candidate problems: DB = two visits of n_scenes scenes, queries = a third visit; every candidate pose that reaches the
    refinement (src scan, tgt scan, constellation transform T) is one problem."""
    cfg, dbc = D.kitti_cm_config(), D.kitti_db_config()
    lb, ub = D.kitti_thres()
    db = c2o.DB(dbc)
    scans = []
    for v in range(2):
        for s0 in range(0, n_scenes, 4):
            pts = synth.make_scans([first_scene + s0 + k for k in range(4)], [v] * 4, n_pts, "cpu", 10 * v + s0).numpy()
            for k in range(4):
                sc = c2o.Scan(cfg, len(scans)).ingest(np.ascontiguousarray(pts[k]))
                db.add_scan(sc, float(len(scans)))
                scans.append(sc)
    for k in range(12):
        db.push_and_balance(k, 5000.0 + k)
    probs = []
    for s0 in range(0, n_scenes, 4):
        pts = synth.make_scans([first_scene + s0 + k for k in range(4)], [2] * 4, n_pts, "cpu", 77 + s0).numpy()
        for k in range(4):
            q = c2o.Scan(cfg, 10000 + s0 + k).ingest(np.ascontiguousarray(pts[k]))
            res, _, _ = db.query(q, lb, ub)
            for c in res["cand"][: min(int(res["n_cand"]), dbc.max_fine_opt)]:
                truth = int(c["cand_gidx"]) % n_scenes == s0 + k  # DB index = visit * n_scenes + scene
                probs.append((scans[int(c["cand_gidx"])], q, np.array(c["T"]), float(c["corr_fine"]), int(c["fine_iters"]), s0 + k, truth))
    return probs


def max_f1_top1(query_ids, corr, truth, n_queries):
    """Max-F1 of the loop-closure decision 'report the query's best candidate if its correlation >= threshold' over all
    thresholds (every query here is a revisit, so recall is counted against n_queries; scripts/pr_mpe.py:71-130 logic)."""
    best = {}
    for qid, c, t in zip(query_ids, corr, truth):
        if qid not in best or c > best[qid][0]:
            best[qid] = (c, t)
    preds = sorted(best.values(), key=lambda x: -x[0])
    tp = fp = 0
    f1 = 0.0
    for c, t in preds:
        tp += bool(t)
        fp += not t
        prec, rec = tp / (tp + fp), tp / n_queries
        if prec + rec > 0:
            f1 = max(f1, 2 * prec * rec / (prec + rec))
    return f1


def refine_study(n_scenes=60, corr_thres=0.65):
    from scipy.optimize import minimize

    probs = refine_problems(n_scenes)
    rows = []
    for a, b, T, corr_fine, iters, qid, truth in probs:
        p0 = np.array([T[2], T[3], np.arctan2(T[1], T[0])])
        r = c2o.refine_solve(a, b, T)
        fun = lambda p: c2o.refine_eval(a, b, T, p)[0]  # noqa: E731
        jac = lambda p: c2o.refine_eval(a, b, T, p)[1]  # noqa: E731
        # the same budget as the reference (max_num_iterations = 10, correlation.h:215) for two independent quasi-Newton codes ...
        lb10 = minimize(fun, p0, jac=jac, method="L-BFGS-B", options=dict(maxiter=10, maxcor=20, ftol=1e-12, gtol=1e-10))
        bf10 = minimize(fun, p0, jac=jac, method="BFGS", options=dict(maxiter=10, gtol=1e-10))
        # ... and the converged optimum as the yardstick
        conv = minimize(fun, p0, jac=jac, method="BFGS", options=dict(maxiter=300, gtol=1e-10))
        norm = r["norm"]
        rows.append((r["correlation"], -lb10.fun / norm, -bf10.fun / norm, -conv.fun / norm, r["iterations"], r["termination"], abs(corr_fine - r["correlation"]), qid, truth))
    R = np.array(rows)
    d_lb, d_bf, d_cv = np.abs(R[:, 0] - R[:, 1]), np.abs(R[:, 0] - R[:, 2]), R[:, 3] - R[:, 0]

    def pct(x):
        return {"median": float(np.median(x)), "p90": float(np.percentile(x, 90)), "p99": float(np.percentile(x, 99)), "max": float(x.max())}

    return {
        "problems": int(len(R)),
        "abs_dcorr_vs_scipy_lbfgsb_10it": pct(d_lb), "abs_dcorr_vs_scipy_bfgs_10it": pct(d_bf),
        "gap_to_converged_optimum": pct(np.maximum(d_cv, 0.0)),
        "restated_solver_never_above_converged": bool((d_cv >= -1e-9).all()),
        "decision_flips_at_thres": {"thres": corr_thres,
                                    "vs_lbfgsb": int(((R[:, 0] >= corr_thres) != (R[:, 1] >= corr_thres)).sum()),
                                    "vs_bfgs": int(((R[:, 0] >= corr_thres) != (R[:, 2] >= corr_thres)).sum()),
                                    "vs_converged": int(((R[:, 0] >= corr_thres) != (R[:, 3] >= corr_thres)).sum())},
        "max_f1_top1_per_query": {"queries": int(n_scenes), "restated": max_f1_top1(R[:, 7], R[:, 0], R[:, 8], n_scenes),
                                  "scipy_lbfgsb_10it": max_f1_top1(R[:, 7], R[:, 1], R[:, 8], n_scenes),
                                  "scipy_bfgs_10it": max_f1_top1(R[:, 7], R[:, 2], R[:, 8], n_scenes),
                                  "converged": max_f1_top1(R[:, 7], R[:, 3], R[:, 8], n_scenes)},
        "iterations_hist": {int(k): int(v) for k, v in zip(*np.unique(R[:, 4], return_counts=True))},
        "query_result_equals_standalone_solve_max_abs": float(R[:, 6].max()),
    }


def layerdb_study(n_scans=4071, n_pts=120000, ts_step=0.104, visits=4):
    """How often the deliberate LayerDB deviation can matter: the reference rebuilds a bucket's KD index only when
    the bucket pops something from its buffer (contour_db.h:119-143); a bucket whose tree a rebalancing move changed
    (contour_db.cpp:63-317) without popping anything is searched through a stale index until its next pop.  Counts, over the
    KITTI-08-shaped sequence of bench.py (same trajectory, same timestamps), the rebalancing moves and the buckets left in
    that state."""
    cfg, dbc = D.kitti_cm_config(), D.kitti_db_config()
    db = c2o.DB(dbc)
    n_scenes = (n_scans + visits - 1) // visits
    for i0 in range(0, n_scans, 8):
        m = min(8, n_scans - i0)
        seeds = [(i0 + k) % n_scenes for k in range(m)]
        vis = [(i0 + k) // n_scenes for k in range(m)]
        pts = synth.make_scans(seeds, vis, n_pts, "cpu", i0).numpy()
        for k in range(m):
            sc = c2o.Scan(cfg, i0 + k).ingest(np.ascontiguousarray(pts[k]))
            db.add_scan(sc, ts_step * (i0 + k))
            db.push_and_balance(i0 + k, ts_step * (i0 + k))
    moves, stale, donors = db.rebalance_stats()
    return {"scans": int(n_scans), "ts_step_s": ts_step, "rebalancing_moves": moves, "buckets_left_with_a_stale_index_in_the_reference": stale,
            "of_which_donor_buckets_undefined_behaviour": donors, "of_which_receiver_buckets_moved_keys_not_yet_searchable": stale - donors}


if __name__ == "__main__":
    import json

    print(json.dumps({"eig": eig_study()}, indent=1))
    print(json.dumps({"refine": refine_study()}, indent=1))
    print(json.dumps({"layerdb": layerdb_study()}, indent=1))
