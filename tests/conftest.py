import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _gpu_usable():
    try:
        import torch
        if not torch.cuda.is_available():
            return False, "no CUDA device"
    except Exception as e:            # pragma: no cover
        return False, "torch unavailable: %s" % e
    lib = os.path.join(REPO, "azul_deep_reinforcement_learning_b200", "libazb.so")
    if not os.path.exists(lib):
        return False, "libazb.so not built (python __graft_entry__.py)"
    return True, ""


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a machine without a CUDA device or without the built library."""
    ok, why = _gpu_usable()
    if ok:
        return
    skip = pytest.mark.skip(reason="gpu test: " + why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
