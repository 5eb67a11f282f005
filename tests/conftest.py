"""Shared fixtures.  `-m "not gpu"` runs here on CPU; `-m gpu` needs a B200 and goes through the C ABI."""
import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "epro-pnp_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def golden_names(prefix=""):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    return {k: g[k] for k in g.files}


def golden_bounds(g, dtype=torch.float32):
    """-> (lb, ub) as None | float | tensor, the way the reference camera holds them."""
    kind = int(g["bounds"])
    if kind == 0:
        return None, None
    if kind == 1:
        return float(g["lb"]), float(g["ub"])
    return torch.from_numpy(g["lb"]).to(dtype), torch.from_numpy(g["ub"]).to(dtype)


def err_vs(a, b):
    """max |a-b| relative to the scale of b (max |b|) -- the "rel" of BASELINE.json's 1e-4."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="session")
def cuda_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
