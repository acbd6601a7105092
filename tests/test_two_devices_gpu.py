"""a process that owns contexts on two GPUs.  Every C-ABI entry point makes its context's device current for the call
and restores the caller's (C2gDeviceGuard), so calls on the two contexts can be interleaved in any order, with either device (or
none of them) current, and give what a single-context run gives.  Skipped on a one-GPU box (run under `gpurun --gpus 2`)."""
import numpy as np
import pytest

from contour_context_b200 import ctypes_defs as D
from helpers import make_batch

pytestmark = pytest.mark.gpu


def test_interleaved_contexts_on_two_devices(built_lib):
    import torch

    from contour_context_b200.engine import Engine

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    lb, ub = D.kitti_thres()
    n_db, n_pts = 12, 40000
    pts, offsets = make_batch([50 + i % 4 for i in range(n_db)], [i // 4 for i in range(n_db)], n_pts, noise_seed=3)
    q, qo = make_batch([50, 51, 52, 53], [3] * 4, n_pts, noise_seed=9)

    def fill(e):
        e.ingest(pts, offsets, first_slot=0, int_ids=np.arange(n_db))
        for i in range(n_db):
            e.db_add_scans(i, 1, [float(i)])
            e.db_push_and_balance(i, float(i))
        for k in range(12):
            e.db_push_and_balance(k, 1000.0 + k)

    torch.cuda.set_device(0)
    ref = Engine(device=0, scan_capacity=32, max_batch=16, max_points=16 * 65536)
    fill(ref)
    ref.ingest(q, qo, first_slot=n_db)
    want = ref.query(n_db, 4, lb, ub).tobytes()
    want_heads = ref.heads(0, n_db).tobytes()
    ref.close()

    a = Engine(device=0, scan_capacity=32, max_batch=16, max_points=16 * 65536)
    torch.cuda.set_device(1)  # the caller's current device is NOT the first context's
    b = Engine(device=1, scan_capacity=32, max_batch=16, max_points=16 * 65536)
    assert torch.cuda.current_device() == 1, "c2g_create must restore the caller's device"
    try:
        # interleave every stage on the two contexts, flipping the current device in between
        a.ingest(pts, offsets, first_slot=0, int_ids=np.arange(n_db))
        torch.cuda.set_device(0)
        b.ingest(pts, offsets, first_slot=0, int_ids=np.arange(n_db))
        for i in range(n_db):
            for e in (b, a):
                e.db_add_scans(i, 1, [float(i)])
                e.db_push_and_balance(i, float(i))
            torch.cuda.set_device(i % 2)
        for k in range(12):
            a.db_push_and_balance(k, 1000.0 + k)
            b.db_push_and_balance(k, 1000.0 + k)
        b.ingest(q, qo, first_slot=n_db)
        a.ingest(q, qo, first_slot=n_db)
        torch.cuda.set_device(1)
        ra = a.query(n_db, 4, lb, ub).tobytes()
        assert torch.cuda.current_device() == 1
        torch.cuda.set_device(0)
        rb = b.query(n_db, 4, lb, ub).tobytes()
        assert a.heads(0, n_db).tobytes() == want_heads and b.heads(0, n_db).tobytes() == want_heads
        assert ra == want and rb == want
    finally:
        a.close()
        b.close()
