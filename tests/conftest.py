import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "footprint-tools_b200")
for p in (PKG, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _have_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    import oracle_lib

    return oracle_lib.load_oracle()


@pytest.fixture(scope="session")
def reflib():
    import oracle_lib

    lib = oracle_lib.load_ref()
    if lib is None:
        pytest.skip("oracle/_ref/libref.so not built (needs /root/reference)")
    return lib


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name), allow_pickle=False)


@pytest.fixture(scope="session")
def ctx():
    from footprint_tools import _native

    return _native.default_context(0)
