"""Geometry helpers on the hot path (reference: quflow/geometry.py)."""
import numpy as np


def hbar(N):
    """Planck-like constant of the quantisation, 2/sqrt(N^2-1) — quflow/geometry.py:7-9."""
    return 2.0 / np.sqrt(float(N) ** 2 - 1.0)
