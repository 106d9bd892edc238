"""TEST INFRASTRUCTURE ONLY — numpy restatement of quflow's mat2shr / shr2mat kernels.

Follows the serial variants of the reference, which state the algebra as matrix-vector products:
``mat2shr_serial_`` (quflow/quantization.py:236-280) and ``shr2mat_serial_`` (:131-169); index helpers
``elm2ind`` (quflow/utils.py:91-105) and ``basis_break_index`` (quantization.py:25-42).  Parity status: PINNED against
outputs of the real reference (oracle/gen_golden_shr.py -> tests/golden/shr_N*.npz, tests/test_oracle.py).
Nothing under ``quflow_b200/`` may import this module.
"""
import numpy as np


def elm2ind(el, m):
    return el * el + el + m                                     # utils.py:105


def basis_break_index(absm, N):
    a = absm - 1                                                # quantization.py:38-41
    return ((a + 2 * a * a - 6 * a * N + 6 * N * N) * (1 + a)) // 6


def _nmax(n_omega, N):
    return N if n_omega >= N * N else int(round(np.sqrt(n_omega)))      # :244-249


def mat2shr(W, basis, n_omega=None):
    N = W.shape[-1]
    omega = np.zeros(N * N if n_omega is None else n_omega)
    Nmax = _nmax(omega.shape[0], N)
    for m in range(Nmax):
        b0 = basis_break_index(m, N)
        Bm = basis[b0:b0 + (N - m) ** 2].reshape(N - m, N - m)
        els = np.arange(m, Nmax)
        if m == 0:
            omega[elm2ind(els, 0)] = ((np.diagonal(W, 0) @ Bm[:, :Nmax]) / 1.0j).real          # :262-264
        else:
            part = np.diagonal(W, -m) @ Bm[:, :Nmax - m]                                       # :268-270
            sgn = 1 if m % 2 == 0 else -1
            omega[elm2ind(els, m)] = np.sqrt(2) * sgn * part.imag                              # :272
            omega[elm2ind(els, -m)] = -np.sqrt(2) * sgn * part.real                            # :276
    return omega / N                                                                            # :278


def shr2mat(omega, basis, N):
    W = np.zeros((N, N), dtype=complex)
    Nmax = _nmax(omega.shape[0], N)
    for m in range(Nmax):
        b0 = basis_break_index(m, N)
        Bm = basis[b0:b0 + (N - m) ** 2].reshape(N - m, N - m)
        els = np.arange(m, Nmax)
        if m == 0:
            d = Bm[:, :Nmax] @ omega[elm2ind(els, 0)].astype(complex)                           # :158-160
            W[np.arange(N), np.arange(N)] = d
        else:
            oc = (1.0 / np.sqrt(2)) * (omega[elm2ind(els, m)] - 1j * omega[elm2ind(els, -m)])   # :163-165
            d = (Bm[:, :Nmax - m] @ oc) * (1 if m % 2 == 0 else -1)                            # :166-168
            i = np.arange(N - m)
            W[i + m, i] = d.conj()                                                              # :169
            W[i, i + m] = d                                                                     # :172
    return W * 1.0j                                                                             # :174
