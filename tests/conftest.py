import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not errored) on a machine without a usable CUDA device."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items:
        return
    try:
        import famseq_b200 as fs

        have = fs.device_count() > 0
    except Exception:  # library not built: let the tests themselves fail loudly
        return
    if not have:
        skip = pytest.mark.skip(reason="no CUDA device visible (run on the GPU box: pytest -m gpu)")
        for it in gpu_items:
            it.add_marker(skip)
