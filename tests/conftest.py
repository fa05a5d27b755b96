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


@pytest.fixture(scope="session")
def c_oracle():
    from oracle import oracle as O
    return O.COracle()
