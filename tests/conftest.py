import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a box without CUDA — a plain `pytest tests` stays green on the CPU.  They are
    NOT skipped when `-m gpu` selects them explicitly on a box that should have a GPU: there a missing device must fail loudly
    (MOBGT_REQUIRE_GPU=1, set by the GPU run scripts)."""
    import torch
    if torch.cuda.is_available() or os.environ.get("MOBGT_REQUIRE_GPU") == "1":
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); there is no CPU fallback for the libmobgt kernels")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def lib_built():
    """Build (or reuse) libmobgt.so; every GPU test calls through it."""
    from mobgt_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "algos_golden.npz"))
