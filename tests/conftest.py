import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """GPU tests are skipped (not failed) when no device is visible, e.g. in the build container."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


@pytest.fixture(scope="session")
def breast_cancer():
    from sklearn.datasets import load_breast_cancer
    from sklearn.preprocessing import StandardScaler
    X, y = load_breast_cancer(return_X_y=True)
    return StandardScaler().fit_transform(X), y


@pytest.fixture(scope="session")
def diabetes():
    from sklearn.datasets import load_diabetes
    from sklearn.preprocessing import StandardScaler
    X, z = load_diabetes(return_X_y=True)
    return StandardScaler().fit_transform(X), z


@pytest.fixture(scope="session")
def digits():
    from sklearn.datasets import load_digits
    from sklearn.preprocessing import StandardScaler
    X, y = load_digits(return_X_y=True)
    return StandardScaler().fit_transform(X), y
