"""Shared pytest plumbing: the ``gpu`` marker and golden-fixture helpers."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
PINNED_CASES = ["debug", "he_small", "large", "be", "ne", "z14"]
MOLECULE_CASES = ["lih", "n2"]
ALL_CASES = PINNED_CASES + MOLECULE_CASES


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Return (system, params, dict of torch tensors) for tests/golden/<name>.npz."""
    from oracle import psiformer_oracle as O

    z = np.load(os.path.join(GOLDEN_DIR, f"{name}.npz"))
    L, H, d, K, nu, nd = (int(v) for v in z["sys_shape"])
    nuclei = tuple((float(zz), tuple(float(c) for c in r)) for zz, r in zip(z["sys_Z"], z["sys_R"]))
    sysm = O.OracleSystem(L, H, d, K, nu, nd, nuclei)
    params = O.synthetic_params(sysm, int(z["param_seed"]))
    data = {}
    for k in z.files:
        a = z[k]
        data[k] = torch.from_numpy(a) if a.dtype.kind in "fbiu" and a.ndim > 0 else a
    return sysm, params, data


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]
    return get
