"""Oracle-side checks of the L-BFGS refinement restatement (oracle/c2o_refine.hpp; Ceres is absent, so parity with the
reference is unpinned — these are the strongest pins available here):
  * the forward-mode dual gradient equals central finite differences of the cost,
  * the solver decreases the cost monotonically from the start point and lands near the optimum an independent
    quasi-Newton implementation (scipy BFGS) finds for the same function,
  * it stops at the reference's iteration cap (max_num_iterations = 10, correlation.h:215)."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth


@pytest.fixture(scope="module")
def scans(oracle):
    cfg = D.kitti_cm_config()
    pts = synth.make_scans([3, 3, 5, 5], [0, 1, 0, 1], 60000, "cpu", 11).numpy()
    return [oracle.Scan(cfg, i).ingest(np.ascontiguousarray(pts[i])) for i in range(4)]


STARTS = ([1.0, 0.0, 0.0, 0.0], [np.cos(0.05), np.sin(0.05), 1.0, -0.5], [np.cos(-0.1), np.sin(-0.1), -2.0, 1.5])


@pytest.mark.parametrize("pair", [(0, 1), (2, 3)])
@pytest.mark.parametrize("T", STARTS)
def test_dual_gradient_matches_finite_differences(oracle, scans, pair, T):
    a, b = scans[pair[0]], scans[pair[1]]
    p0 = np.array([T[2], T[3], np.arctan2(T[1], T[0])])
    cost, grad, n_pairs = oracle.refine_eval(a, b, T, p0)
    assert n_pairs > 50 and np.isfinite(cost) and cost < 0
    for k in range(3):
        h = 1e-6
        pp, pm = p0.copy(), p0.copy()
        pp[k] += h
        pm[k] -= h
        fd = (oracle.refine_eval(a, b, T, pp)[0] - oracle.refine_eval(a, b, T, pm)[0]) / (2 * h)
        assert abs(fd - grad[k]) <= 1e-5 * max(1.0, abs(grad[k])), (k, fd, grad[k])


@pytest.mark.parametrize("pair", [(0, 1), (2, 3)])
@pytest.mark.parametrize("T", STARTS)
def test_solver_against_scipy_bfgs(oracle, scans, pair, T):
    from scipy.optimize import minimize

    a, b = scans[pair[0]], scans[pair[1]]
    p0 = np.array([T[2], T[3], np.arctan2(T[1], T[0])])
    r = oracle.refine_solve(a, b, T)
    assert r["termination"] in (0, 1) and 1 <= r["iterations"] <= 10
    assert r["final_cost"] <= r["initial_cost"]
    assert abs(r["correlation"] - (-r["final_cost"] / r["norm"])) < 1e-12
    # the returned parameters reproduce the returned cost
    assert abs(oracle.refine_eval(a, b, T, r["x"])[0] - r["final_cost"]) <= 1e-9 * abs(r["final_cost"])
    sp = minimize(lambda p: oracle.refine_eval(a, b, T, p)[0], p0, jac=lambda p: oracle.refine_eval(a, b, T, p)[1], method="BFGS",
                  options=dict(gtol=1e-10, maxiter=300))
    # 10 iterations of L-BFGS recover most of what a converged quasi-Newton run gains (never more than it)
    gain_ours, gain_ref = r["initial_cost"] - r["final_cost"], r["initial_cost"] - sp.fun
    assert gain_ours <= gain_ref + 1e-6 * abs(sp.fun)
    assert gain_ours >= 0.5 * gain_ref, (gain_ours, gain_ref)
    if r["termination"] == 1:
        assert abs(r["final_cost"] - sp.fun) <= 1e-4 * abs(sp.fun)
