import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def built_lib():
    """libbppgpu.so, built on demand (nvcc cross-compiles without a GPU)."""
    from bpp_phyl_b200 import capi
    if not capi.LIB_PATH.exists():
        import __graft_entry__ as g
        g.build()
    return capi.lib()
