"""Shared helpers for the parity tests (tests/ only)."""
import numpy as np

from contour_context_b200 import ctypes_defs as D
from contour_context_b200 import synth


def make_batch(scene_seeds, visits, n_pts, noise_seed=0):
    """numpy float32 [B, n_pts, 4] + int64 offsets [B+1]."""
    pts = synth.make_scans(scene_seeds, visits, n_pts, device="cpu", noise_seed=noise_seed).numpy()
    B = pts.shape[0]
    offsets = np.arange(B + 1, dtype=np.int64) * n_pts
    return np.ascontiguousarray(pts.reshape(-1, 4)), offsets


def view_fields_equal(a: np.ndarray, b: np.ndarray):
    """Bitwise comparison of two c2g_view arrays, field by field; returns list of mismatching field names."""
    bad = []
    for name in D.VIEW_DTYPE.names:
        if name == "pad_":
            continue
        if a[name].tobytes() != b[name].tobytes():
            bad.append(name)
    return bad


def f32_bits(x):
    return np.ascontiguousarray(x, np.float32).view(np.uint32)


def ulp_diff(a, b):
    a = f32_bits(a).astype(np.int64)
    b = f32_bits(b).astype(np.int64)
    a = np.where(a < 0x80000000, a, 0x80000000 - a)
    b = np.where(b < 0x80000000, b, 0x80000000 - b)
    return np.abs(a - b)
