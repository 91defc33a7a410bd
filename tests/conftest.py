import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "timeout: per-test time limit (pytest-timeout; ignored when the plugin is absent)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no CUDA device is visible, e.g. a plain `pytest tests/`."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        # a GPU test that hangs (a device-side wait, a third-party library that never answers) must fail, not stall the run
        for it in items:
            if "gpu" in it.keywords and not any(m.name == "timeout" for m in it.iter_markers()):
                it.add_marker(pytest.mark.timeout(300))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the CPU-side libraries exist (cheap no-op when already built)."""
    import __graft_entry__ as g
    g.build_cpu()
