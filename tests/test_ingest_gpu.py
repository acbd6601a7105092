"""GPU parity of the ingest path (BEV -> contours -> views -> keys -> BCI) against the CPU oracle, through the C-ABI.

Bar (BASELINE.json north_star): bit-exact BEV cells, contour order, ContourView fields, retrieval keys and BCIs.  The libm
calls that reach keys / BCIs (exp, atan2f) run as bit-exact glibc restatements on the device (csrc/c2g_libm.cuh,
tests/test_libm.py); if the host's exp() is not one of the two glibc variants the context reports exp_mode 0 and the key
comparison falls back to a counted <= 2 ulp bound."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from helpers import make_batch, ulp_diff, view_fields_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine(built_lib):
    from contour_context_b200.engine import Engine

    e = Engine(scan_capacity=64, max_batch=16, max_points=16 * 131072)
    yield e
    e.close()


def _oracle_scans(oracle, cfg, pts, offsets):
    out = []
    for b in range(len(offsets) - 1):
        out.append(oracle.Scan(cfg, b).ingest(pts[offsets[b]:offsets[b + 1]]))
    return out


@pytest.mark.parametrize("n_pts", [120000, 30001])
def test_bev_bit_exact(engine, oracle, n_pts):
    pts, offsets = make_batch([3, 3, 4, 5], [0, 1, 0, 2], n_pts)
    engine.ingest(pts, offsets, first_slot=0)
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        ob, orf, ocf = s.bev()
        gb, grf, gcf = engine.bev(b)
        assert gb.tobytes() == ob.tobytes(), f"scan {b}: bev heights differ"
        assert grf.tobytes() == orf.tobytes() and gcf.tobytes() == ocf.tobytes(), f"scan {b}: pillar coords differ"


def _compact_from_dense(cfg, bev, rf, cf):
    """What the scatter kernel must hand to the contour kernel, derived from the oracle's dense BEV."""
    n_row, n_col = cfg.n_row, cfg.n_col
    wpr = (n_col + 31) // 32
    h = bev.reshape(n_row, n_col)
    planes = np.zeros((D.NLEV, n_row, wpr), np.uint32)
    for lev in range(D.NLEV):
        rows, cols = np.nonzero(h > np.float32(cfg.lv_grads[lev]))
        np.bitwise_or.at(planes[lev], (rows, cols >> 5), (np.uint32(1) << (cols & 31).astype(np.uint32)))
    fgm = (h > np.float32(cfg.lv_grads[0])).reshape(-1)
    fg = np.stack([bev[fgm], rf[fgm], cf[fgm], np.zeros(int(fgm.sum()), np.float32)], axis=1).astype(np.float32)
    return planes, fg, int((bev > -999.0).sum())


def test_scatter_handoff_planes_and_foreground_list(engine, oracle):
    """K1's output (bit-planes, raster-ordered foreground list, occupied count) against the oracle's BEV, for the production
    variant (low points only mark occupancy) and the full-tile variant behind c2g_get_bev; ragged / NaN / tie inputs included."""
    base, _ = make_batch([20, 23], [0, 1], 70000)
    scans = [base[:70000], base[70000:], base[:3000].copy(), base[:9000].copy(), base[:17].copy()]
    scans[2][::5, 0] = np.nan
    scans[3][:, 2] = np.round(scans[3][:, 2] * 2) / 2    # many exact height ties -> first point in file order wins
    scans[4][:, :2] = 500.0                               # empty BEV
    pts = np.ascontiguousarray(np.concatenate(scans))
    offsets = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    engine.ingest_bev_only(pts, offsets)
    n_fg_total = 0
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        ob, orf, ocf = s.bev()
        planes, fg, n_occ = _compact_from_dense(engine.cm_cfg, ob, orf, ocf)
        for full in (False, True):
            gp, gf, gocc = engine.bev_compact(b, full_tile_variant=full)
            assert gocc == n_occ, (b, full, gocc, n_occ)
            assert gp.tobytes() == planes.tobytes(), f"scan {b} full={full}: bit-planes differ"
            assert gf.shape == fg.shape and gf.tobytes() == fg.tobytes(), f"scan {b} full={full}: foreground list differs"
        n_fg_total += len(fg)
    assert n_fg_total > 2000


def test_scatter_fast_path_and_its_deferrals(engine, oracle):
    """The 32-bit scatter kernel logs (height, cell, index) events and resolves the winners afterwards; scans it cannot take go to
    the 64-bit kernel: more than 2^17 points (index field), more events than the log holds (heights ascending in file order:
    every point above the lowest threshold is a record), more foreground cells than the rank table (cluttered test below).
    Mixed in one batch with ordinary scans; everything must match the oracle, and exactly the two odd scans must be deferred."""
    base, _ = make_batch([31, 32, 33], [0, 1, 2], 70000)
    big = np.concatenate([base[:70000], base[70000:140000]])                 # 140 000 points > 2^17
    rng = np.random.default_rng(11)
    asc = np.zeros((60000, 4), np.float32)                                   # 900 cells, heights ascending in file order:
    asc[:, :2] = rng.uniform(10.0, 40.0, (60000, 2))                         # ~every point raises its cell -> ~60 000 events
    asc[:, 2] = np.linspace(0.0, 6.0, 60000)
    ties = base[:60000].copy()
    ties[:, 2] = np.round(ties[:, 2] * 4) / 4                                # exact ties: the first point in file order wins
    scans = [base[:70000], big, asc, ties, base[70000:140000]]
    pts = np.ascontiguousarray(np.concatenate(scans))
    offsets = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    engine.ingest_bev_only(pts, offsets)
    assert engine.scatter_deferred() == 2
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        ob, orf, ocf = s.bev()
        planes, fg, n_occ = _compact_from_dense(engine.cm_cfg, ob, orf, ocf)
        gp, gf, gocc = engine.bev_compact(b)
        assert gocc == n_occ, (b, gocc, n_occ)
        assert gp.tobytes() == planes.tobytes(), f"scan {b}: bit-planes differ"
        assert gf.shape == fg.shape and gf.tobytes() == fg.tobytes(), f"scan {b}: foreground list differs"
    engine.ingest_bev_only(pts[: offsets[1]], offsets[:2])
    assert engine.scatter_deferred() == 0


def test_views_keys_bci(engine, oracle):
    seeds, visits = [10, 10, 11, 12, 13, 13], [0, 1, 0, 0, 2, 3]
    pts, offsets = make_batch(seeds, visits, 120000)
    engine.ingest(pts, offsets, first_slot=8, int_ids=np.arange(100, 106))
    heads = engine.heads(8, len(seeds))
    key_ulp_bad = 0
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        oh = s.head()
        gh = heads[b]
        assert gh["status"] == 0
        assert gh["int_id"] == 100 + b
        assert np.array_equal(gh["n_views"], oh["n_views"]), (b, gh["n_views"], oh["n_views"])
        assert np.array_equal(gh["layer_cell_cnt"], oh["layer_cell_cnt"])
        assert gh["n_occupied"] == oh["n_occupied"]
        assert np.array_equal(gh["n_ell"], oh["n_ell"])
        gviews = engine.views(8 + b, gh)
        for lev in range(D.NLEV):
            ov = s.views(lev)
            bad = view_fields_equal(gviews[lev], ov)
            assert not bad, f"scan {b} level {lev}: view fields differ: {bad}"
        # keys: bit-exact except for counted <= 2-ulp libm effects
        gk, ok = gh["keys"], oh["keys"]
        nan_g, nan_o = np.isnan(gk), np.isnan(ok)
        assert np.array_equal(nan_g, nan_o)
        d = ulp_diff(np.where(nan_g, 0, gk), np.where(nan_o, 0, ok))
        assert d.max() <= (2 if engine.exp_mode() == 0 else 0), f"scan {b}: key differs by {d.max()} ulp"
        key_ulp_bad += int((d > 0).sum())
        # BCI: bitsets, neighbour identity and order, segments exact; r exact; theta within 1 ulp (atan2f libm caveat)
        gb_, ob_ = gh["bcis"], oh["bcis"]
        assert gb_["dist_bin"].tobytes() == ob_["dist_bin"].tobytes()
        for f in ("n_nei", "n_seg", "piv_seq", "level", "seg"):
            assert np.array_equal(gb_[f], ob_[f]), f
        for f in ("level", "seq", "bit_pos"):
            assert np.array_equal(gb_["nei"][f], ob_["nei"][f]), f
        assert gb_["nei"]["r"].tobytes() == ob_["nei"]["r"].tobytes()
        assert gb_["nei"]["theta"].tobytes() == ob_["nei"]["theta"].tobytes()  # glibc atan2f restated on the device
        assert abs(gh["gmm_auto_corr"] - oh["gmm_auto_corr"]) <= 1e-9 * abs(oh["gmm_auto_corr"])
    # at most a handful of 1-ulp key entries over 6 scans x 360 key entries
    assert key_ulp_bad <= 4, f"{key_ulp_bad} key entries differ from the oracle"


def test_ragged_and_degenerate_inputs(engine, oracle):
    """Ragged batch: different point counts per scan, an (almost) empty scan, points outside the BEV, NaNs, ties."""
    rng = np.random.default_rng(7)
    base, _ = make_batch([20], [0], 50000)
    scans = [base[:50000], base[:1000], base[:17].copy(), base[:5000].copy(), base[:8000].copy()]
    scans[2][:, :2] = 500.0                      # everything outside the square -> empty BEV
    scans[3][::7, 0] = np.nan                    # NaN x dropped
    scans[3][::11, 2] = np.nan                   # NaN z never stored
    scans[4][:, 2] = np.round(scans[4][:, 2])    # many exact height ties -> first point in file order wins
    pts = np.ascontiguousarray(np.concatenate(scans))
    offsets = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    engine.ingest(pts, offsets, first_slot=0)
    heads = engine.heads(0, len(scans))
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        ob, orf, ocf = s.bev()
        gb, grf, gcf = engine.bev(b)
        assert gb.tobytes() == ob.tobytes() and grf.tobytes() == orf.tobytes() and gcf.tobytes() == ocf.tobytes(), b
        oh = s.head()
        assert np.array_equal(heads[b]["n_views"], oh["n_views"])
        gviews = engine.views(b, heads[b])
        for lev in range(D.NLEV):
            assert not view_fields_equal(gviews[lev], s.views(lev)), (b, lev)
        gk, ok = heads[b]["keys"], oh["keys"]
        assert np.array_equal(np.isnan(gk), np.isnan(ok))
        assert ulp_diff(np.nan_to_num(gk), np.nan_to_num(ok)).max() <= (2 if engine.exp_mode() == 0 else 0)


def test_device_resident_input_matches_host_input(engine):
    import torch

    pts, offsets = make_batch([30, 31], [0, 1], 60000)
    engine.ingest(pts, offsets, first_slot=0)
    h_host = engine.heads(0, 2).copy()
    t = torch.from_numpy(pts).cuda()
    engine.ingest(t, offsets, first_slot=2)
    h_dev = engine.heads(2, 2)
    for f in ("n_views", "layer_cell_cnt", "keys", "n_ell"):
        assert h_host[f].tobytes() == h_dev[f].tobytes(), f


def test_xyz_input_variant_matches_xyzi(engine):
    """c2g_ingest_xyz (12 B per point, intensity dropped on the host) must give byte-identical descriptors, from host and device
    buffers, ragged batches included, and the dense-image getter must work on such a batch."""
    import torch

    base, _ = make_batch([40, 41, 42], [0, 1, 2], 50000)
    scans = [base[:50000], base[50000:87001], base[100000:100017]]
    pts = np.ascontiguousarray(np.concatenate(scans))
    offsets = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    xyz = np.ascontiguousarray(pts[:, :3])
    engine.ingest(pts, offsets, first_slot=0, int_ids=np.arange(3))
    ref_h = engine.heads(0, 3).copy()
    ref_v = [np.concatenate(engine.views(b, ref_h[b])).tobytes() for b in range(3)]
    ref_bev = [engine.bev(b) for b in range(3)]
    for dev in (False, True):
        src = torch.from_numpy(xyz).cuda() if dev else xyz
        engine.ingest_xyz(src, offsets, first_slot=4, int_ids=np.arange(3))
        h = engine.heads(4, 3)
        assert h.tobytes() == ref_h.tobytes(), f"device={dev}: heads differ"
        for b in range(3):
            assert np.concatenate(engine.views(4 + b, h[b])).tobytes() == ref_v[b]
            got = engine.bev(b)
            assert all(a.tobytes() == r.tobytes() for a, r in zip(got, ref_bev[b]))


def _cells_to_points(h, rng):
    """One point per occupied cell (jittered inside the cell), z so that lidar_height + z == h."""
    rows, cols = np.nonzero(h > -999.0)
    n = len(rows)
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = rows - 75 + rng.uniform(0.1, 0.9, n)
    pts[:, 1] = cols - 75 + rng.uniform(0.1, 0.9, n)
    pts[:, 2] = h[rows, cols] - 2.0
    return pts[rng.permutation(n)]


def test_cluttered_scans_use_the_overflow_arenas(engine, oracle):
    """Salt-and-pepper BEVs: ~5 000 runs per level (run tables spill to the global arena), one 16 000-cell component, and a
    second scan with ~4 500 components (component tables spill too).  Everything must still match the oracle bit for bit."""
    rng = np.random.default_rng(5)
    scans = []
    for (a, b, frac, cmax) in [(-1.0, 9.0, 1.0, 150), (0.5, 8.5, 0.45, 85), (1.0, 6.0, 0.12, 150)]:
        h = np.full((150, 150), -1000.0)
        m = rng.random((150, 150)) < frac
        m[:, cmax:] = False
        h[m] = rng.uniform(a, b, int(m.sum()))
        scans.append(_cells_to_points(h, rng))
    base, _ = make_batch([21], [0], 60000)
    scans.append(base)                            # an ordinary scan in the same batch (shared-memory tables)
    pts = np.ascontiguousarray(np.concatenate(scans))
    offsets = np.cumsum([0] + [len(s) for s in scans]).astype(np.int64)
    engine.ingest(pts, offsets, first_slot=0)
    heads = engine.heads(0, len(scans))
    for b, s in enumerate(_oracle_scans(oracle, engine.cm_cfg, pts, offsets)):
        oh = s.head()
        assert heads[b]["status"] == 0, (b, heads[b]["status"], oh["n_views"])
        assert np.array_equal(heads[b]["n_views"], oh["n_views"]), (b, heads[b]["n_views"], oh["n_views"])
        assert np.array_equal(heads[b]["layer_cell_cnt"], oh["layer_cell_cnt"])
        gviews = engine.views(b, heads[b])
        for lev in range(D.NLEV):
            assert not view_fields_equal(gviews[lev], s.views(lev)), (b, lev)
        gk, ok = heads[b]["keys"], oh["keys"]
        assert np.array_equal(np.isnan(gk), np.isnan(ok))
        assert ulp_diff(np.nan_to_num(gk), np.nan_to_num(ok)).max() <= (2 if engine.exp_mode() == 0 else 0)
        assert heads[b]["bcis"]["dist_bin"].tobytes() == oh["bcis"]["dist_bin"].tobytes()
