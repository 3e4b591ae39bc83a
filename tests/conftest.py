import os
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "hermespy-rt_b200"))
sys.path.insert(0, os.path.dirname(__file__))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "ref: needs oracle/_ref (the compiled reference)")


@pytest.fixture(scope="session")
def root():
    return ROOT


def _gpu_usable() -> bool:
    try:
        import hrt_b200
        return hrt_b200.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not errored) where no CUDA device is usable --
    a plain `pytest tests` in the GPU-less authoring container stays green."""
    if not any("gpu" in it.keywords for it in items) or _gpu_usable():
        return
    skip = pytest.mark.skip(reason="no usable CUDA device (the product has no CPU path)")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
