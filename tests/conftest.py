import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # GPU tests fail loudly rather than skip when selected with -m gpu on a box without a GPU;
    # without -m they are skipped on CPU-only machines.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no GPU in this container (run with -m gpu on the B200 box)")
        for item in items:
            if "gpu" in item.keywords:
                item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import fos_oracle
    fos_oracle.build()
    return fos_oracle


@pytest.fixture(scope="session")
def fos():
    import fos_b200
    return fos_b200


@pytest.fixture(scope="session")
def gpu_handle_factory(fos):
    def make(**options):
        H = fos.Handle(0)
        for k, v in options.items():
            H.set_option(k, v)
        return H
    return make
