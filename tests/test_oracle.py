"""CPU tests: pin the oracle (oracle/) against the reference.

Sources of truth, in order of strength:
 1. outputs of the real reference run in the build container, stored by
    oracle/gen_golden.py under tests/golden (the reference tree itself does
    not exist on the GPU box);
 2. the reference's own known answers: its N=16 golden vector
    (tests/test_integrators.py:58-319 upstream), exact Poisson solutions on
    the spherical-harmonic basis (tests/test_laplacian.py:226-252 upstream,
    atol 1e-14*N^2), the Laplacian eigenvalue test (:134-152), and the
    notebook's printed norms of the N=512 initial condition;
 3. cross-checks between independent restatements (C vs numpy Thomas),
    and a live comparison against /root/reference when it is present.
"""
import numpy as np
import pytest

import oracle
from oracle import refshim
from conftest import golden, unband, relfro


# ----------------------------------------------------------------- Poisson
@pytest.mark.parametrize("name,N", [("poisson_exact_N33_zt1", 33), ("poisson_exact_N33_zt0", 33),
                                    ("poisson_exact_N64_zt1", 64), ("poisson_exact_N64_zt0", 64),
                                    ("poisson_exact_N101_zt1", 101)])
def test_solve_poisson_exact_and_reference(name, N):
    g = golden(name + ".npz")
    P = oracle.solve_poisson(g["Wexact"])
    # upstream tests/test_laplacian.py:252
    np.testing.assert_allclose(P, g["Pexact"], atol=1e-14 * N ** 2, rtol=0)
    # the reference's own output on the same input (rounding-level agreement)
    assert relfro(P, g["P_ref"]) < 1e-13
    # exactly skew-Hermitian and trace-free like the reference's construction
    assert np.abs(P + P.conj().T).max() == 0.0
    assert abs(np.trace(P)) < 1e-12


@pytest.mark.parametrize("name,N", [("poisson_exact_N33_zt1", 33), ("poisson_exact_N64_zt1", 64),
                                    ("poisson_exact_N101_zt1", 101)])
def test_laplace_eigen(name, N):
    g = golden(name + ".npz")
    W = oracle.laplace(g["Pexact"])
    np.testing.assert_allclose(W, g["Wexact"], rtol=1e-7, atol=1e-9)   # upstream :152 (default rtol)
    assert relfro(W, g["W_ref"]) < 1e-14


@pytest.mark.parametrize("N", [64, 127])
def test_solve_poisson_random(N):
    g = golden(f"poisson_random_N{N}.npz")
    P = oracle.solve_poisson(g["W"])
    assert relfro(P, g["P_ref"]) < 1e-13
    P2 = oracle.solve_poisson_numpy(g["W"])
    assert relfro(P2, g["P_ref"]) < 1e-13
    # round trip: Δ Δ^{-1} W = W - tr(W)/N
    W = g["W"]
    assert relfro(oracle.laplace(P), W - np.eye(N) * np.trace(W) / N) < 1e-11


def test_laplacian_table_matches_reference():
    g = golden("poisson_random_N64.npz")
    assert np.array_equal(oracle.laplacian(64, bc=True), g["lap"])
    assert np.array_equal(oracle.isomp_oracle.laplacian_numpy(64, bc=True), g["lap"])


def test_hbar():
    assert oracle.hbar(512) == 2.0 / np.sqrt(512 ** 2 - 1)


@pytest.mark.parametrize("N", [2, 3, 5, 16])
def test_tiny_sizes(N):
    W = oracle.random_skewherm(N, 1)
    P = oracle.solve_poisson(W)
    assert relfro(oracle.solve_poisson_numpy(W), P) < 1e-13
    assert relfro(oracle.laplace(P), W) < 1e-12


# ------------------------------------------------------------------- isomp
def test_reference_golden_vector_N16():
    """The reference's own golden W0 -> Wfinal (500 steps, stepsize 0.02).

    It predates the trace-removal / bc change of the default solve_poisson
    (SURVEY.md §4), so it is replayed with the legacy Poisson semantics; this
    pins the integrator loop itself.  With HEAD semantics the oracle must match
    the reference's HEAD output instead.
    """
    g = golden("ref_isomp_golden_N16.npz")
    dt = oracle.hbar(16) * float(g["stepsize"])
    steps = int(g["steps"])
    W = oracle.isomp(g["W0"].copy(), dt, steps, hamiltonian=lambda X: oracle.solve_poisson(X, legacy=True))
    np.testing.assert_allclose(W, g["Wfinal"], rtol=0, atol=1e-7)      # upstream :45
    assert relfro(W, g["W_legacy"]) < 1e-12
    W = oracle.isomp(g["W0"].copy(), dt, steps)
    assert relfro(W, g["W_head"]) < 1e-12


@pytest.mark.parametrize("N", [32, 64, 128])
def test_isomp_random_matches_reference(N):
    g = golden(f"isomp_R_N{N}.npz")
    W0 = oracle.random_skewherm(N, 42)
    assert np.array_equal(W0, g["W0"])
    steps = int(g["steps"]) if N < 128 else 100
    rec, stats = {}, {'iterations': 0.0}
    W = oracle.isomp(W0.copy(), float(g["dt"]), steps, stats=stats, record=rec)
    assert stats['tol_auto'] == pytest.approx(float(g["tol_auto"]), rel=1e-14)
    assert rec['iterations'] == list(g["iterations"][:steps])
    ref = g["Wfinal"] if N < 128 else g["W_step100"]
    assert relfro(W, ref) < 1e-12
    if N < 128:
        assert stats['iterations'] == float(g["mean_iterations"])
        np.testing.assert_allclose(oracle.casimirs(W), g["casimirs"], rtol=1e-10)


@pytest.mark.parametrize("tag,kw", [("compsum", dict(compsum=True)), ("tol1e-10", dict(tol=1e-10)),
                                    ("reinit", dict(reinitialize=True)), ("minit3", dict(minit=3, maxit=5))])
def test_isomp_option_variants(tag, kw):
    g = golden(f"isomp_R_N32_{tag}.npz")
    rec, stats = {}, {'iterations': 0.0}
    W = oracle.isomp(oracle.random_skewherm(32, 42), float(g["dt"]), int(g["steps"]), stats=stats, record=rec, **kw)
    assert rec['iterations'] == list(g["iterations"])
    assert stats['number_of_maxit'] == float(g["number_of_maxit"])
    assert relfro(W, g["Wfinal"]) < 1e-12


def test_isomp_profile_mode():
    g = golden("isomp_R_N64_profile.npz")
    rec = {}
    W = oracle.isomp(oracle.random_skewherm(64, 42), float(g["dt"]), 5, minit=10, maxit=10, record=rec)
    assert rec['iterations'] == [10] * 5
    assert relfro(W, g["Wfinal"]) < 1e-13


def test_isomp_smooth_N64():
    g = golden("isomp_S_N64.npz")
    W0 = unband(g["W0_band"], 64)
    rec = {}
    W = oracle.isomp(W0.copy(), float(g["dt"]), 100, record=rec)
    assert rec['iterations'] == list(g["iterations"])
    assert relfro(W, g["Wfinal"]) < 1e-12


def test_isomp_smooth_N512_known_answers():
    """Config 2 input: notebook cell 7 prints L2 = 0.9999999999999999 and
    Linf (spectral) = 2.6779890041828724 for this initial condition."""
    g = golden("isomp_S_N512.npz")
    N = 512
    W0 = unband(g["W0_band"], N)
    assert np.linalg.norm(W0) / np.sqrt(N) == pytest.approx(0.9999999999999999, rel=1e-14)
    assert np.linalg.norm(W0, 2) == pytest.approx(2.6779890041828724, rel=1e-11)
    rec = {}
    W = oracle.isomp(W0.copy(), float(g["dt"]), 100, record=rec)
    assert rec['iterations'] == list(g["iterations"])
    idx = g["sample_idx"]
    scale = float(g["normF"]) / N
    assert np.abs(W[idx[:, 0], idx[:, 1]] - g["Wfinal_sample"]).max() < 1e-11 * max(scale, np.abs(g["Wfinal_sample"]).max())
    assert np.abs(W[:48, :48] - g["Wfinal_block"]).max() < 1e-12 * np.abs(g["Wfinal_block"]).max()
    assert np.linalg.norm(W) == pytest.approx(float(g["normF"]), rel=1e-13)


@pytest.mark.parametrize("case", ["callback", "forcing", "forcing_time", "strang", "ham_scaled", "ham_time", "all"])
def test_isomp_hooks(case):
    """callback / forcing / strang_splitting / custom and time-dependent Hamiltonians against the reference's own
    output (oracle/gen_golden.py part H)."""
    from oracle import hooks
    g = golden("isomp_hooks_N32.npz")
    kw = hooks.case_kwargs(case, oracle.solve_poisson)
    rec, stats, cb = {}, {'iterations': 0.0}, []
    W = oracle.isomp(g["W0"].copy(), float(g["dt"]), int(g["steps"]), stats=stats, record=rec,
                     callback=lambda W, dW: cb.append((np.linalg.norm(W), np.linalg.norm(dW))), **kw)
    assert rec['iterations'] == list(g[f"{case}_iterations"])
    assert stats['tol_auto'] == pytest.approx(float(g[f"{case}_tol_auto"]), rel=1e-14)
    assert relfro(W, g[f"{case}_Wfinal"]) < 1e-12
    np.testing.assert_allclose(np.array(cb), g[f"{case}_cb_norms"], rtol=1e-11)
    if case in ("forcing", "forcing_time"):
        with pytest.raises(NotImplementedError):      # isospectral.py:588-589
            oracle.isomp(g["W0"].copy(), float(g["dt"]), 1, compsum=True, **kw)


@pytest.mark.parametrize("tag,kw", [("plain", dict()), ("compsum", dict(compsum=True))])
def test_isomp_multistate(tag, kw):
    """(k, N, N) input: members 1.. are advected by member 0's stream function, tolerance and residual come from
    member 0 (cpu.py:672-674, isospectral.py:444-446, 528-531)."""
    g = golden("isomp_multistate_N32.npz")
    rec, stats = {}, {'iterations': 0.0}
    W = oracle.isomp(g["W0"].copy(), float(g["dt"]), int(g["steps"]), stats=stats, record=rec, **kw)
    assert rec['iterations'] == list(g[f"{tag}_iterations"])
    assert stats['tol_auto'] == pytest.approx(float(g[f"{tag}_tol_auto"]), rel=1e-14)
    assert relfro(W, g[f"{tag}_Wfinal"]) < 1e-12


def test_asserts_and_stats_semantics():
    W = oracle.random_skewherm(8, 3)
    with pytest.raises(AssertionError):
        oracle.isomp(W.copy(), 0.1, 1, minit=0)
    with pytest.raises(AssertionError):
        oracle.isomp(W.copy(), 0.1, 1, minit=3, maxit=2)
    empty = {}
    oracle.isomp(W.copy(), 0.01, 2, stats=empty)        # `if stats:` — an empty dict is ignored (isospectral.py:451,609)
    assert empty == {}
    out = W.copy()
    ret = oracle.isomp(out, 0.01, 2)
    assert ret is out                                    # updated in place and returned


# ----------------------------------------------- live check (build container only)
@pytest.mark.skipif(not refshim.available(), reason="reference tree only exists in the build container")
def test_live_against_reference():
    refshim.load()
    from quflow.integrators.isospectral import isomp_fixedpoint
    from quflow.laplacian.cpu import solve_poisson, laplace
    N = 48
    W0 = oracle.random_skewherm(N, 5)
    assert relfro(oracle.solve_poisson(W0), solve_poisson(W0).copy()) < 1e-13
    assert relfro(oracle.laplace(W0), laplace(W0)) < 1e-14
    dt = 0.3 * oracle.hbar(N)
    st_ref, st = {'iterations': 0.0}, {'iterations': 0.0}
    Wref = isomp_fixedpoint(W0.copy(), dt, steps=40, stats=st_ref)
    W = oracle.isomp(W0.copy(), dt, steps=40, stats=st)
    assert st == st_ref or (st['iterations'] == st_ref['iterations'] and st['tol_auto'] == pytest.approx(st_ref['tol_auto'], rel=1e-14))
    assert relfro(W, Wref) < 1e-12


@pytest.mark.parametrize("N", [16, 33])
def test_shr_oracle_against_reference_golden(N):
    """The mat2shr / shr2mat restatement against outputs of the real reference (oracle/gen_golden_shr.py)."""
    from oracle import shr_oracle
    g = golden(f"shr_N{N}.npz")
    scale = np.abs(g["omega_full"]).max()
    assert np.abs(shr_oracle.mat2shr(g["W"], g["basis"]) - g["omega_full"]).max() < 1e-13 * scale
    assert np.abs(shr_oracle.mat2shr(g["W"], g["basis"], 64) - g["omega_trunc"]).max() < 1e-13 * scale
    assert relfro(shr_oracle.shr2mat(g["omega_band"], g["basis"], N), g["W_band"]) < 1e-13
    assert relfro(shr_oracle.shr2mat(g["omega_N"], g["basis"], N), g["W_N"]) < 1e-13
