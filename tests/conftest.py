"""Shared pytest configuration: `gpu` marker, import paths, golden-vector loader, parity metric."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "ideal-gan_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]
    return load


def rel_err(x, ref):
    """Parity metric of BASELINE.md §4: max|x - ref| / max|ref| (per tensor)."""
    x = np.asarray(x, dtype=np.complex128 if np.iscomplexobj(x) or np.iscomplexobj(ref) else np.float64)
    ref = np.asarray(ref, dtype=x.dtype)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    denom = np.max(np.abs(ref))
    if denom == 0:
        return float(np.max(np.abs(x)))
    return float(np.max(np.abs(x - ref)) / denom)


def assert_close(x, ref, tol, what=""):
    e = rel_err(x, ref)
    assert e <= tol, f"{what}: rel-to-max error {e:.3e} > {tol:.1e}"
