"""GPU parity tests for the operators on the hot path: solve_poisson, laplace, norm, ZGEMM.

Every call goes through the C ABI (ctypes -> libquflow_b200.so).  Tolerances are floating-point
tolerances, stated per test; the oracle is the CPU restatement pinned in tests/test_oracle.py.
"""
import numpy as np
import pytest

import oracle
from conftest import golden, relfro

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def qf(cuda_device):
    import quflow_b200
    return quflow_b200


def to_dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


@pytest.mark.parametrize("name,N", [("poisson_exact_N33_zt1", 33), ("poisson_exact_N33_zt0", 33),
                                    ("poisson_exact_N64_zt1", 64), ("poisson_exact_N64_zt0", 64),
                                    ("poisson_exact_N101_zt1", 101)])
def test_solve_poisson_golden(qf, name, N):
    g = golden(name + ".npz")
    P = qf.solve_poisson(g["Wexact"])            # host path
    np.testing.assert_allclose(P, g["Pexact"], atol=1e-14 * N ** 2, rtol=0)   # upstream tests/test_laplacian.py:252
    assert relfro(P, g["P_ref"]) < 1e-13          # vs the reference's own output
    Pd = qf.solve_poisson(to_dev(g["Wexact"])).cpu().numpy()                   # device path
    assert np.array_equal(Pd, P)
    assert np.abs(P + P.conj().T).max() == 0.0
    assert abs(np.trace(P)) < 1e-12


# 2050 / 3000: bands longer than one CTA -> thread-block clusters of 2 with DSMEM carry exchange; 5000: clusters of 4 that
# hold several linked groups
@pytest.mark.parametrize("N", [2, 3, 5, 16, 31, 32, 33, 64, 127, 128, 129, 200, 257, 512, 1100, 2048, 2050, 3000, 5000])
def test_solve_poisson_vs_oracle(qf, N):
    W = oracle.random_skewherm(N, seed=N)
    P = qf.solve_poisson(W)
    Pref = oracle.solve_poisson(W)
    assert relfro(P, Pref) < 1e-13
    # only the upper triangle of W is read (reference semantics): garbage below the diagonal is ignored
    W2 = W.copy()
    W2[np.tril_indices(N, -1)] = 7.0 + 3.0j
    assert np.array_equal(qf.solve_poisson(W2), P)


def test_solve_poisson_nonzero_trace_and_roundtrip(qf):
    N = 96
    W = oracle.random_skewherm(N, 3) + 0.3j * np.eye(N)
    P = qf.solve_poisson(W)
    assert relfro(P, oracle.solve_poisson(W)) < 1e-13
    back = qf.laplace(P)
    assert relfro(back, W - np.eye(N) * np.trace(W) / N) < 1e-11


def test_solve_poisson_multistate_and_time_kwarg(qf):
    N = 40
    W = np.stack([oracle.random_skewherm(N, 1), oracle.random_skewherm(N, 2)])
    P = qf.solve_poisson(W)                        # reduce=select_first, cpu.py:696-697
    assert P.shape == (N, N)
    assert np.array_equal(P, qf.solve_poisson(W[0]))
    with pytest.raises(TypeError):                 # isomp's autonomy probe relies on this (isospectral.py:416-423)
        qf.solve_poisson(W[0], time=0.0)


@pytest.mark.parametrize("N", [2, 33, 65, 128, 300])
def test_laplace(qf, N):
    rng = np.random.RandomState(N)
    P = rng.randn(N, N) + 1j * rng.randn(N, N)     # general (not skew-Hermitian) input
    assert relfro(qf.laplace(P), oracle.laplace(P)) < 1e-15
    if N in (33,):
        g = golden("poisson_exact_N33_zt1.npz")
        np.testing.assert_allclose(qf.laplace(g["Pexact"]), g["Wexact"], rtol=1e-7, atol=1e-9)   # upstream :152


@pytest.mark.parametrize("N", [7, 64, 129, 500])
def test_norm_inf(qf, N):
    from quflow_b200._cuda import get_handle
    W = oracle.random_skewherm(N, 11)
    got = get_handle(N).norm_inf(to_dev(W))[0]
    assert got == pytest.approx(np.linalg.norm(W, np.inf), rel=1e-14)


@pytest.mark.parametrize("N", [8, 33, 64, 100, 128, 129, 192, 256, 333, 512])
def test_zgemm_vs_numpy(qf, N):
    """Complex FP64 DMMA GEMM against numpy (BLAS zgemm): both are fp64 dot products of length N, so the
    difference is bounded by accumulation-order rounding, ~ sqrt(N) * eps relative to |A||B| (the default 3M
    arithmetic forms the imaginary part as T3 - T1 - T2, which roughly doubles the constant)."""
    from quflow_b200._cuda import get_handle
    rng = np.random.RandomState(N)
    A = rng.randn(N, N) + 1j * rng.randn(N, N)
    B = rng.randn(N, N) + 1j * rng.randn(N, N)
    C = get_handle(N).zgemm(to_dev(A), to_dev(B)).cpu().numpy()
    ref = A @ B
    bound = np.abs(A) @ np.abs(B)
    assert (np.abs(C - ref) / bound).max() < 1e-15 * np.sqrt(N) + 2e-15
    # exactness on small integers: any summation order gives the same result
    Ai = rng.randint(-3, 4, (N, N)) + 1j * rng.randint(-3, 4, (N, N))
    Bi = rng.randint(-3, 4, (N, N)) + 1j * rng.randint(-3, 4, (N, N))
    Ci = get_handle(N).zgemm(to_dev(Ai.astype(complex)), to_dev(Bi.astype(complex))).cpu().numpy()
    assert np.array_equal(Ci, Ai @ Bi)


def test_zgemm_batched(qf):
    from quflow_b200._cuda import get_handle
    N, k = 96, 3
    rng = np.random.RandomState(0)
    A = rng.randn(k, N, N) + 1j * rng.randn(k, N, N)
    B = rng.randn(k, N, N) + 1j * rng.randn(k, N, N)
    C = get_handle(N, k).zgemm(to_dev(A), to_dev(B)).cpu().numpy()
    assert relfro(C, A @ B) < 1e-14


def test_inner_products_and_loggers(qf):
    """quflow/geometry.py:53-76 and quflow/physics.py:9-38 on the device: inner_L2, norm_L2, energy_euler, enstrophy,
    H^-1 / H^1 inner products.  Tolerance 1e-13 relative (only the summation order differs from numpy)."""
    import torch
    N = 96
    W = oracle.random_skewherm(N, 4)
    V = oracle.random_skewherm(N, 5)
    P = oracle.solve_poisson(W)
    ref_inner = float((V * W.conj()).sum().real / N)
    assert qf.inner_L2(V, W) == pytest.approx(ref_inner, rel=1e-13)
    assert qf.norm_L2(W) == pytest.approx(np.linalg.norm(W) / np.sqrt(N), rel=1e-14)
    assert qf.enstrophy(W) == pytest.approx(float((W * W.conj()).sum().real / N) / 2.0, rel=1e-14)
    e_ref = -float((W * P.conj()).sum().real / N) / 2.0
    assert qf.energy_euler(W) == pytest.approx(e_ref, rel=1e-12)
    assert qf.energy_euler(torch.from_numpy(W).to("cuda:0")) == qf.energy_euler(W)      # deterministic, same path
    assert qf.inner_Hm1(V, W) == pytest.approx(-float((V * P.conj()).sum().real / N), rel=1e-11)
    assert qf.norm_Hm1(W) == pytest.approx(np.sqrt(2.0 * e_ref), rel=1e-12)
    LW = oracle.laplace(P)
    assert qf.inner_H1(V, P) == pytest.approx(-float((V * LW.conj()).sum().real / N), rel=1e-11)
    assert qf.norm_H1(P) == pytest.approx(np.sqrt(-float((P * LW.conj()).sum().real / N)), rel=1e-12)
    # energy and enstrophy are invariants of the flow: conserved by isomp to the fixed-point tolerance
    W1 = W.copy()
    qf.isomp(W1, 0.25 * qf.hbar(N), steps=20)
    assert qf.enstrophy(W1) == pytest.approx(qf.enstrophy(W), rel=1e-9)
    assert qf.energy_euler(W1) == pytest.approx(qf.energy_euler(W), rel=1e-8)      # near-conserved (oracle: -7e-10)


@pytest.mark.parametrize("N", [16, 33])
def test_mat2shr_shr2mat_vs_reference_golden(N):
    """mat2shr / shr2mat on the device against the REAL reference's outputs for its own basis
    (oracle/gen_golden_shr.py -> tests/golden/shr_N*.npz; reference: quantization.py:172-227, 283-325)."""
    import torch
    import quflow_b200 as qf
    from quflow_b200.quantization import device_basis, basis_size
    g = golden(f"shr_N{N}.npz")
    assert basis_size(N) == g["basis"].size
    B = device_basis(g["basis"])
    scale = np.abs(g["omega_full"]).max()
    om = qf.mat2shr(g["W"], B)                                   # numpy in -> numpy out
    assert isinstance(om, np.ndarray) and np.abs(om - g["omega_full"]).max() < 1e-13 * scale
    om7 = qf.mat2shr(torch.from_numpy(g["W"]).cuda(), B, elmax=7)    # device in -> device out, truncated
    assert om7.is_cuda and np.abs(om7.cpu().numpy() - g["omega_trunc"]).max() < 1e-13 * scale
    Wb = qf.shr2mat(g["omega_band"], g["basis"], N=N)            # basis as numpy: uploaded by the call
    assert np.abs(Wb - g["W_band"]).max() < 1e-13 * np.abs(g["W_band"]).max()
    assert np.abs(Wb + Wb.conj().T).max() == 0.0                 # exactly skew-Hermitian
    WN = qf.shr2mat(g["omega_N"], B)
    assert relfro(WN, g["W_N"]) < 1e-13
    # round trip through the device: mat2shr(shr2mat(omega)) == omega
    back = qf.mat2shr(qf.shr2mat(torch.from_numpy(g["omega_N"]).cuda(), B), B)
    assert np.abs(back.cpu().numpy() - g["omega_N"]).max() < 1e-12 * np.abs(g["omega_N"]).max()
