"""CPU-only: the C-ABI library loads and exports every symbol include/c2g.h declares; struct sizes agree."""
import ctypes as C
import os
import re

from contour_context_b200 import ctypes_defs as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header(built_lib):
    hdr = open(os.path.join(ROOT, "include", "c2g.h")).read()
    names = set(re.findall(r"\b(c2g_[a-z_0-9]+)\s*\(", hdr))
    assert len(names) >= 15
    for n in sorted(names):
        assert hasattr(built_lib, n), f"libc2g.so does not export {n}"


def test_struct_sizes(built_lib):
    assert built_lib.c2g_abi_version() == 1
    assert built_lib.c2g_sizeof(0) == D.SCAN_HEAD_DTYPE.itemsize
    assert built_lib.c2g_sizeof(1) == D.VIEW_DTYPE.itemsize == 80
    assert built_lib.c2g_sizeof(2) == D.BCI_DTYPE.itemsize == 608
    assert built_lib.c2g_sizeof(3) == D.HINT_DTYPE.itemsize == 16
    assert built_lib.c2g_sizeof(4) == D.PAIR_SCORE_DTYPE.itemsize == 128
    assert built_lib.c2g_sizeof(5) == D.QUERY_RESULT_DTYPE.itemsize
    assert built_lib.c2g_sizeof(6) == C.sizeof(D.CmConfig)
    assert built_lib.c2g_sizeof(7) == C.sizeof(D.DbConfig)


def test_no_oracle_in_product():
    """The product package must never import / link the oracle."""
    pkg = os.path.join(ROOT, "contour_context_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, fn), errors="ignore").read()
                assert "oracle" not in src.replace("// oracle note", "").replace("see oracle note", ""), f"{fn} mentions the oracle"
