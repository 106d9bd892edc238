import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def unband(band, N):
    """Inverse of oracle/gen_golden.py:band() — rebuild a banded matrix."""
    lmax = (band.shape[0] - 1) // 2
    W = np.zeros((N, N), dtype=band.dtype)
    for m in range(-lmax, lmax + 1):
        n = N - abs(m)
        W += np.diag(band[m + lmax, :n], m)
    return W


def relfro(A, B):
    return float(np.linalg.norm(A - B) / np.linalg.norm(B))


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
