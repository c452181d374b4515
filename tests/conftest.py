import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """a plain `pytest tests` on a box without a CUDA device (or without the built library) skips the GPU tests
    instead of failing them; `-m gpu` on the B200 box runs them"""
    import torch

    reason = None
    if not torch.cuda.is_available():
        reason = "no CUDA device"
    elif not os.path.exists(os.path.join(ROOT, "wsovod_b200", "libwsovod_b200.so")):
        reason = "libwsovod_b200.so not built (python -m wsovod_b200.build)"
    if reason:
        skip = pytest.mark.skip(reason=reason)
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import torch

    def load(name):
        return torch.load(os.path.join(ROOT, "tests", "golden", name + ".pt"), weights_only=False)

    return load
