import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built_lib():
    """libc2g.so must exist (built in-tree by __graft_entry__.build()); build it if the sources are newer."""
    import __graft_entry__ as g

    g.build_c2g()
    from contour_context_b200 import capi

    return capi.lib()


@pytest.fixture(scope="session")
def oracle():
    from oracle import c2o

    c2o.build()
    return c2o
