"""TEST INFRASTRUCTURE ONLY — numpy/C restatement of quflow's isomp hot path.

This module is the parity oracle for the CUDA path.  It restates, on the CPU,

* ``hbar``                      — quflow/geometry.py:7-9
* ``laplacian`` table           — quflow/laplacian/cpu.py:55-95, 604-625
* ``solve_poisson`` / Thomas    — quflow/laplacian/cpu.py:281-362, 681-734
* ``laplace``                   — quflow/laplacian/cpu.py:98-108, 628-669
* ``conj_subtract_``            — quflow/integrators/isospectral.py:66-81
* ``isomp_fixedpoint``          — quflow/integrators/isospectral.py:338-613

It must never be imported from ``quflow_b200`` (the product path has no CPU
fallback).  The heavy loops live in ``poisson_oracle.c`` (gcc, OpenMP over
diagonals like the reference's numba ``prange``); the matrix products go
through ``np.matmul`` (BLAS zgemm) exactly like the reference
(isospectral.py:496,499).  A second, dependency-free numpy implementation of
the Thomas solve (vectorised across diagonals) is kept as a cross-check.

Parity status: PINNED.  ``oracle/gen_golden.py`` imports the real reference
in the build container and stores its outputs under ``tests/golden``;
``tests/test_oracle.py`` checks this module against those fixtures and
against the reference's own known answers.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


# ---------------------------------------------------------------------------
# C half
# ---------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> str:
    """Compile ``poisson_oracle.c`` into ``oracle/libqforacle.so`` (gcc)."""
    so = os.path.join(_HERE, "libqforacle.so")
    src = os.path.join(_HERE, "poisson_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B", "libqforacle.so"])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build_c_oracle())
        p = ctypes.c_void_p
        lib.qfo_compute_laplacian.argtypes = [ctypes.c_int, ctypes.c_double, p]
        lib.qfo_compute_laplacian.restype = None
        lib.qfo_solve_poisson_skewh.argtypes = [ctypes.c_int, p, p, p, p, p, ctypes.c_int]
        lib.qfo_solve_poisson_skewh.restype = None
        lib.qfo_laplace.argtypes = [ctypes.c_int, p, p, p]
        lib.qfo_laplace.restype = None
        lib.qfo_conj_subtract.argtypes = [ctypes.c_int, p, p]
        lib.qfo_conj_subtract.restype = None
        lib.qfo_norm_inf.argtypes = [ctypes.c_int, p]
        lib.qfo_norm_inf.restype = ctypes.c_double
        _LIB = lib
    return _LIB


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(ctypes.c_void_p)


# ---------------------------------------------------------------------------
# geometry / laplacian
# ---------------------------------------------------------------------------
def hbar(N: int) -> float:
    """quflow/geometry.py:7-9"""
    return 2.0 / np.sqrt(float(N) ** 2 - 1.0)


_lap_cache: dict = {}


def laplacian(N: int, bc: bool = False, bc_shift: float = -0.5) -> np.ndarray:
    """(N,N,2) coefficient table, quflow/laplacian/cpu.py:55-95 (cached like :604-625)."""
    key = (N, bc, bc_shift)
    if key not in _lap_cache:
        lap = np.zeros((N, N, 2), dtype=np.float64)
        _lib().qfo_compute_laplacian(N, bc_shift if bc else 0.0, _ptr(lap))
        _lap_cache[key] = lap
    return _lap_cache[key]


def laplacian_numpy(N: int, bc: bool = False, bc_shift: float = -0.5) -> np.ndarray:
    """Same table, pure numpy (cross-check of the C code)."""
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    m = np.abs(j - i).astype(np.float64)
    k = np.minimum(i, j).astype(np.float64)
    lap = np.zeros((N, N, 2))
    lap[..., 0] = -((N - 1.0) * (2.0 * k + 1.0 + m) - 2.0 * k * (k + m))
    lap[..., 1] = np.sqrt(((k + m) * (N - k - m)) * (k * (N - k)))
    if bc:
        lap[0, 0, 0] += bc_shift
    return lap


def solve_poisson(W: np.ndarray, legacy: bool = False) -> np.ndarray:
    """P = Δ_N^{-1} W,  quflow/laplacian/cpu.py:681-734 (dense branch).

    Returns a fresh array (the reference returns a module-level cached buffer,
    cpu.py:726 — callers here never rely on that aliasing).
    ``legacy=True`` = quflow/laplacian/gpu.py:73,143-173 semantics (bc +0.5, no trace
    removal of W, trace removal of P), needed only to replay the reference's stale N=16 golden vector.
    """
    if W.ndim >= 3:  # reduce=select_first, cpu.py:672-674,696-697
        W = np.ascontiguousarray(W[(0,) * (W.ndim - 2) + (Ellipsis,)])
    W = np.ascontiguousarray(W, dtype=np.complex128)
    N = W.shape[-1]
    lap = laplacian(N, bc=True, bc_shift=0.5 if legacy else -0.5)
    P = np.zeros_like(W)
    bf = np.zeros((N, N), dtype=np.float64)
    bcx = np.zeros((N, N), dtype=np.complex128)
    _lib().qfo_solve_poisson_skewh(N, _ptr(lap), _ptr(W), _ptr(P), _ptr(bf), _ptr(bcx), 2 if legacy else 3)
    return P


def solve_poisson_numpy(W: np.ndarray) -> np.ndarray:
    """Second opinion: the same Thomas sweeps, vectorised across diagonals in numpy.

    Element (k, k+m) belongs to system m at position k, so row k of the upper
    triangle holds position k of systems m = 0..N-1-k.
    """
    W = np.asarray(W, dtype=np.complex128)
    N = W.shape[-1]
    lap = laplacian_numpy(N, bc=True)
    d, o = lap[..., 0], lap[..., 1]
    trW = np.trace(W) / N
    u = np.zeros((N, N))
    c = np.zeros((N, N), dtype=np.complex128)
    # position 0 of every system is row 0
    u[0, :] = d[0, :]
    c[0, :] = W[0, :]
    c[0, 0] -= trW
    for k in range(1, N):
        sl = slice(k, N)            # columns j = k+m, m = 0..N-1-k
        pl = slice(k - 1, N - 1)    # previous position: (k-1, j-1)
        w = o[k, sl] / u[k - 1, pl]
        u[k, sl] = d[k, sl] - w * o[k, sl]
        c[k, sl] = W[k, sl] - w * c[k - 1, pl]
        c[k, k] -= trW
    P = np.zeros((N, N), dtype=np.complex128)
    # last position of system m is row N-1-m, column N-1
    for k in range(N - 1, -1, -1):
        sl = slice(k, N)
        x = c[k, sl] / u[k, sl]
        if k < N - 1:
            # systems with a successor: columns k..N-2 -> successor (k+1, j+1)
            nxt = o[k + 1, k + 1:N] * P[k + 1, k + 1:N]
            x[:-1] = (c[k, k:N - 1] - nxt) / u[k, k:N - 1]
        P[k, sl] = x
    iu = np.triu_indices(N, 1)
    P[iu[1], iu[0]] = -np.conj(P[iu])
    P[np.diag_indices(N)] -= np.trace(P) / N
    return P


def laplace(P: np.ndarray) -> np.ndarray:
    """W = Δ_N P, quflow/laplacian/cpu.py:628-669 (dense branch) / :98-108."""
    P = np.ascontiguousarray(P, dtype=np.complex128)
    N = P.shape[-1]
    lap = laplacian(N, bc=False)
    W = np.zeros_like(P)
    _lib().qfo_laplace(N, _ptr(lap), _ptr(P), _ptr(W))
    return W


def conj_subtract_(a: np.ndarray, out: np.ndarray) -> None:
    """out = a - a^H with exact skew-symmetry, isospectral.py:66-81 (2-D branch, and member by member for 3-D)."""
    assert a.flags.c_contiguous and out.flags.c_contiguous
    if a.ndim == 3:                                          # :75-81
        for k in range(a.shape[0]):
            _lib().qfo_conj_subtract(a.shape[-1], _ptr(a[k]), _ptr(out[k]))
        return
    assert a.ndim == 2
    _lib().qfo_conj_subtract(a.shape[-1], _ptr(a), _ptr(out))


def norm_inf(A: np.ndarray) -> float:
    """max row sum of |z| (np.linalg.norm(A, inf) / scipy.linalg.norm(A, ord=inf))."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    return float(_lib().qfo_norm_inf(A.shape[-1], _ptr(A)))


# ---------------------------------------------------------------------------
# the integrator
# ---------------------------------------------------------------------------
def isomp_fixedpoint(W, dt, steps=100, hamiltonian=None, time=None, forcing=None,
                     strang_splitting=None, stats=None, callback=None, tol='auto',
                     maxit=10, minit=1, verbatim=False, compsum=False, reinitialize=False,
                     record=None):
    """Isospectral midpoint by fixed-point iteration, isospectral.py:338-613.

    Restated for a 2-D state and for the multi-state (k, N, N) mode (members 1.. are advected passively by member 0's
    stream function: select_first, cpu.py:672-674; tolerance and residual from member 0, isospectral.py:444-446, 529-531),
    including the callers' hooks of the loop: ``callback`` (:550-551),
    ``forcing`` (:403-414, :511-520, :594-596), ``strang_splitting`` (:466-467, :602-603) and custom or
    time-dependent Hamiltonians (:416-424, :488-491).
    ``record`` (oracle-only extra): dict that receives per-step ``iterations``
    and ``resnorm`` lists so tests can compare iteration counts step by step.
    """
    assert minit >= 1, "minit must be at least 1."          # :400
    assert maxit >= minit, "maxit must be at minit."         # :401
    if hamiltonian is None:
        hamiltonian = solve_poisson
    assert W.ndim in (2, 3)

    if forcing is not None:                                  # :403-414
        autonomous_force = True
        if time is not None:
            try:
                FW = forcing(W, W, time=time)
            except TypeError:
                pass
            else:
                autonomous_force = False
        FW = np.zeros_like(W)
    autonomous = True                                        # :416-424
    if time is not None:
        try:
            Phalf = hamiltonian(W, time=time)
        except TypeError:
            pass
        else:
            autonomous = False

    total_iterations = 0                                     # :426-427
    number_of_maxit = 0
    dW = np.zeros_like(W)                                    # :430-433
    dW_old = np.zeros_like(W)
    Whalf = np.zeros_like(W)
    PWcomm = np.zeros_like(W)
    hb = hbar(W.shape[-1])                                   # :436-437
    vareps = dt / (2 * hb)

    if (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0):   # :440-452
        mach_eps = np.finfo(W.dtype).eps
        if not compsum:
            mach_eps = np.sqrt(mach_eps)
        tol = (mach_eps * dt / hb) * np.linalg.norm(W[0] if W.ndim > 2 else W, np.inf)   # :444-448
        if verbatim:
            print("Tolerance set to {}.".format(tol))
        if stats:
            stats['tol_auto'] = tol

    if compsum:                                              # :455-459
        y_compsum = np.zeros_like(W)
        c_compsum = np.zeros_like(W)
        t_compsum = np.zeros_like(W)
        delta_compsum = np.zeros_like(W)

    if record is not None:
        record.setdefault('iterations', [])
        record.setdefault('resnorm', [])
        record['tol'] = float(tol)

    for k in range(steps):                                   # :463
        if strang_splitting:                                 # :466-467
            W = strang_splitting(dt / 2, W)
        resnorm = np.inf                                     # :470-472
        if reinitialize:
            dW.fill(0.0)
        its = 0
        for i in range(maxit):                               # :475
            total_iterations += 1
            its += 1
            np.copyto(Whalf, W)                              # :481-482
            Whalf += dW
            np.copyto(dW_old, dW)                            # :485
            if autonomous:                                   # :488-491
                Phalf = hamiltonian(Whalf)
            else:
                Phalf = hamiltonian(Whalf, time=time + dt / 2)
            Phalf *= vareps                                  # :492
            np.matmul(Phalf, Whalf, out=PWcomm)              # :496
            np.matmul(PWcomm, Phalf, out=dW)                 # :499
            conj_subtract_(PWcomm, PWcomm)                   # :503
            dW += PWcomm                                     # :509
            if forcing:                                      # :511-520
                Phalf /= vareps
                if autonomous_force:
                    FW = forcing(Phalf, Whalf)
                else:
                    FW = forcing(Phalf, Whalf, time=time + dt / 2)
                FW *= dt / 2
                dW += FW
            if i + 1 >= minit:                               # :523-536
                resnorm_old = resnorm
                dW_old -= dW
                if dW_old.ndim > 2:
                    # scipy.linalg.norm(., ord=inf, axis=(-1, -2)) (:528): row axis -1, column axis -2, i.e. the max
                    # COLUMN sum of each member; member 0 decides when the Hamiltonian returns one matrix (:529-530)
                    resnorm = float(np.abs(dW_old[0]).sum(axis=-2).max())
                    if not np.isfinite(np.abs(dW_old).sum()):
                        resnorm = np.nan
                else:
                    resnorm = norm_inf(dW_old)
                if not np.isfinite(resnorm):                 # scipy's asarray_chkfinite
                    raise ValueError("array must not contain infs or NaNs")
                if resnorm <= tol or resnorm >= resnorm_old:
                    break
        else:                                                # :538-542
            number_of_maxit += 1
            if verbatim:
                print("Max iterations {} reached at step {}.".format(maxit, k))
        if record is not None:
            record['iterations'].append(its)
            record['resnorm'].append(float(resnorm))

        PWcomm *= 2                                          # :547
        if callback is not None:                             # :550-551
            callback(W, PWcomm)
        if compsum:                                          # :553-586 (Kahan)
            np.copyto(y_compsum, PWcomm)
            y_compsum -= c_compsum
            np.copyto(t_compsum, W)
            t_compsum += y_compsum
            np.copyto(delta_compsum, t_compsum)
            delta_compsum -= W
            np.copyto(c_compsum, delta_compsum)
            c_compsum -= y_compsum
            np.copyto(W, t_compsum)
            if forcing:                                      # :588-589
                raise NotImplementedError("Compensated sum with forcing is not yet implemented.")
        else:
            W += PWcomm                                      # :592
            if forcing:                                      # :594-596
                FW *= 2
                W += FW
        if time is not None:                                 # :598-599
            time += dt
        if strang_splitting:                                 # :602-603
            W = strang_splitting(dt / 2, W)

    if verbatim:
        print("Average number of iterations per step: {:.2f}".format(total_iterations / steps))
    if stats:                                                # :609-611
        stats["iterations"] = total_iterations / steps
        stats["number_of_maxit"] = number_of_maxit / steps
    return W


isomp = isomp_fixedpoint                                     # :617


# ---------------------------------------------------------------------------
# helpers shared by tests / bench (inputs and invariants)
# ---------------------------------------------------------------------------
def random_skewherm(N: int, seed: int = 42) -> np.ndarray:
    """R(N, seed): tests' get_random_mat (reference tests/test_integrators.py:12-18)
    normalised to norm_L2 = ||W||_F / sqrt(N) = 1."""
    rng = np.random.RandomState(seed)
    A = rng.randn(N, N) + 1j * rng.randn(N, N)
    W = A - A.conj().T
    W -= np.eye(N) * np.trace(W) / N
    W /= np.linalg.norm(W) / np.sqrt(N)
    return np.ascontiguousarray(W)


def casimirs(W: np.ndarray, kmax: int = 4) -> np.ndarray:
    """C_k(W) = tr((iW)^k)/N for k = 2..kmax (real for skew-Hermitian W)."""
    H = 1j * W
    lam = np.linalg.eigvalsh((H + H.conj().T) / 2)
    return np.array([np.sum(lam ** k) / W.shape[-1] for k in range(2, kmax + 1)])
