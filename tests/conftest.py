"""pytest configuration: registers the `gpu` marker and shared helpers.

`-m "not gpu"` = oracle vs golden vectors, host logic, C-ABI symbol checks (no CUDA calls).
`-m gpu`       = parity tests proper: CUDA path (through the C-ABI) vs the oracle.
"""
import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, prefix + "*.npz")))


def load_golden(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """max |a-b| / max |b|  -- the norm-wise relative error every parity bound here is stated in."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.size == 0 and b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def l2_err(a, b):
    """||a-b||_2 / ||b||_2 -- insensitive to where the largest entry sits."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def per_layer_err(a, b):
    """max over the leading (layer) axis of max|a_l - b_l| / max|b_l|: a small layer's gradient cannot
    hide behind a large one's (layers whose reference gradient is identically zero are skipped)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    worst = 0.0
    for l in range(b.shape[0]):
        m = np.abs(b[l]).max()
        if m > 0:
            worst = max(worst, float(np.abs(a[l] - b[l]).max() / m))
    return worst


def grad_errs(o, ref, suffix=""):
    """Every metric the tensor-core bounds are stated in, for d_ws / d_bs."""
    out = {}
    for k in ("d_ws", "d_bs"):
        r = ref[k + suffix]
        out[k] = rel_err(o[k], r)
        out[k + "_layer"] = per_layer_err(o[k], r)
        out[k + "_l2"] = l2_err(o[k], r)
    return out


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import oracle as O
    return O.COracle()
