import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def native_lib():
    """Build (if stale) and load libexacto_b200.so."""
    import __graft_entry__ as g
    g.build()
    from exacto_b200 import _native
    return _native.lib()
