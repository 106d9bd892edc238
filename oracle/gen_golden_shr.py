#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — fixtures for mat2shr / shr2mat from the REAL reference (build container only).

    python oracle/gen_golden_shr.py   ->  tests/golden/shr_N{16,33}.npz

Each fixture holds the reference's quantization basis for that N (quflow.quantization.compute_basis, flat layout,
quantization.py:25-42), a random skew-Hermitian W with omega = mat2shr(W) (full and truncated to elmax = 7), and a random
band-limited omega with W = shr2mat(omega, N)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from oracle.isomp_oracle import random_skewherm  # noqa: E402

refshim.load(with_quantization=True)
import quflow.quantization as quq  # noqa: E402


def main():
    for N in (16, 33):
        basis = quq.compute_basis(N)
        W = random_skewherm(N, 7)
        omega_full = np.zeros(N * N)
        quq.mat2shr_(W, basis, omega_full)
        omega_trunc = np.zeros(8 * 8)
        quq.mat2shr_(W, basis, omega_trunc)
        rng = np.random.RandomState(3)
        om = np.zeros(6 * 6)
        om[1:] = rng.randn(35)
        Wb = np.zeros((N, N), dtype=complex)
        quq.shr2mat_(om, basis, Wb)
        omN = rng.randn(N * N)
        omN[0] = 0.0
        WN = np.zeros((N, N), dtype=complex)
        quq.shr2mat_(omN, basis, WN)
        out = os.path.join(ROOT, "tests", "golden", f"shr_N{N}.npz")
        np.savez_compressed(out, basis=basis, W=W, omega_full=omega_full, omega_trunc=omega_trunc, omega_band=om, W_band=Wb,
                            omega_N=omN, W_N=WN)
        print("wrote", out, os.path.getsize(out) // 1024, "KiB", "basis", basis.shape)


if __name__ == "__main__":
    main()
