import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "scalable-bayesian-matrix-factorization_b200")
GOLDEN = os.path.join(ROOT, "tests", "golden")
for p in (PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    if os.environ.get("SBMF_EMULATED"):   # tools/sbmf_sanitize.sh: the GPU tests against the host build of the kernels (SBMF_LIB_PATH)
        return True
    try:
        import ctypes
        cudart = ctypes.CDLL("libcudart.so")  # noqa: F841
    except OSError:
        pass
    return os.path.exists("/dev/nvidia0") or os.path.exists("/dev/nvidiactl")


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no GPU in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def ml100k():
    d = np.load(os.path.join(GOLDEN, "ml100k.npz"))
    out = {k: d[k].astype(np.uint32 if "rating" not in k else np.float32) for k in d.files}
    out["num_users"] = int(max(out["train_user"].max(), out["test_user"].max())) + 1
    out["num_items"] = int(max(out["train_item"].max(), out["test_item"].max())) + 1
    return out


@pytest.fixture(scope="session")
def ml100k_split():
    """the reference's second real fixture, data/m1m/m100k (80k/20k split, item ids not re-based: items 0..942 are empty rows)"""
    d = np.load(os.path.join(GOLDEN, "ml100k_split.npz"))
    out = {k: d[k].astype(np.uint32 if "rating" not in k else np.float32) for k in d.files}
    out["num_users"] = int(max(out["train_user"].max(), out["test_user"].max())) + 1
    out["num_items"] = int(max(out["train_item"].max(), out["test_item"].max())) + 1
    return out


def _triples(path):
    a = np.loadtxt(path, dtype=np.float64, ndmin=2)
    return a[:, 0].astype(np.uint32), a[:, 1].astype(np.uint32), a[:, 2].astype(np.float32)


@pytest.fixture(scope="session")
def tiny():
    tu, ti, tr = _triples(os.path.join(GOLDEN, "tiny_unsorted.train"))
    su, si, sr = _triples(os.path.join(GOLDEN, "tiny_unsorted.test"))
    return {"train_user": tu, "train_item": ti, "train_rating": tr, "test_user": su, "test_item": si, "test_rating": sr,
            "num_users": int(max(tu.max(), su.max())) + 1, "num_items": int(max(ti.max(), si.max())) + 1}
