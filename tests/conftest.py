import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """A kernel that never returns would block the GPU box until the harness kills the call: every GPU
    test gets a hard deadline (pytest-timeout's thread method ends the process even inside a blocked
    CUDA call)."""
    if not config.pluginmanager.hasplugin("timeout"):
        return
    for item in items:
        if item.get_closest_marker("gpu") and not item.get_closest_marker("timeout"):
            item.add_marker(pytest.mark.timeout(240, method="thread"))


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the checker libs (and the CUDA lib if missing) once per session."""
    import __graft_entry__ as g
    g.build_cuda()      # no-op when libwrcu.so is newer than its sources
    g.build_host()
    g.build_oracle()
