"""Isospectral midpoint integrator on the B200 — drop-in for ``quflow.integrators.isomp``.

Reference: quflow/integrators/isospectral.py:338-613 (``isomp_fixedpoint``; alias ``isomp`` :617).
The signature, argument meaning, in-place semantics, ``stats`` keys and error behaviour mirror the
reference.  The whole fixed-point loop runs inside the CUDA library (csrc/isomp.cu) without host
synchronisation; Python only validates arguments and converts the statistics.
"""
import numpy as np

from ._cuda import get_handle
from .geometry import hbar
from .laplacian import solve_poisson, _is_torch, _prepare


def _check_hamiltonian(hamiltonian):
    """The device path implements the default Hamiltonian only; it is recognised by identity, like the
    reference's own GPU seam (quflow/simulation.py:554)."""
    if hamiltonian is None or hamiltonian is solve_poisson:
        return
    mod = getattr(hamiltonian, "__module__", "") or ""
    if getattr(hamiltonian, "__name__", "") == "solve_poisson" and mod.startswith("quflow.laplacian"):
        return   # the reference's own default, when this package is plugged in behind quflow
    raise NotImplementedError("quflow_b200.isomp runs the default Hamiltonian (solve_poisson) on the device; "
                              "custom Hamiltonians are outside the accelerated hot path")


def isomp_fixedpoint(W,
                     dt,
                     steps=100,
                     hamiltonian=solve_poisson,
                     time=None,
                     forcing=None,
                     strang_splitting=None,
                     stats=None,
                     callback=None,
                     tol='auto',
                     maxit=10,
                     minit=1,
                     verbatim=False,
                     compsum=False,
                     reinitialize=False
                     ):
    """Time-stepping by the isospectral midpoint method with fixed-point iterations
    (reference: isospectral.py:338-613).

    ``W``: skew-Hermitian (N, N) complex128; a numpy array is overwritten in place and returned, exactly like
    the reference (host→device and back once per call); a torch CUDA tensor is advanced in place on the device.
    ``time`` is accepted (``qf.solve`` always passes it, simulation.py:727) — the default Hamiltonian is
    autonomous so it has no effect.  ``forcing``, ``strang_splitting`` and ``callback`` are not part of the
    accelerated path and raise NotImplementedError (as the reference's own GPU prototype does,
    experimental/isospectral_cuda.py:191,332).
    """
    assert minit >= 1, "minit must be at least 1."          # isospectral.py:400
    assert maxit >= minit, "maxit must be at minit."         # isospectral.py:401
    _check_hamiltonian(hamiltonian)
    if forcing is not None:
        raise NotImplementedError("forcing is not implemented on the device path")
    if strang_splitting is not None:
        raise NotImplementedError("strang_splitting is not implemented on the device path")
    if callback is not None:
        raise NotImplementedError("callback is not implemented on the device path")
    if W.ndim != 2:
        raise NotImplementedError("multi-state (k, N, N) input is not implemented; use quflow_b200.isomp_ensemble "
                                  "for independent members")
    Wc = _prepare(W)
    inplace = Wc is W or (_is_torch(W) and Wc.data_ptr() == W.data_ptr())
    N = Wc.shape[-1]
    auto = (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0)   # :440
    if isinstance(tol, str) and not auto:
        raise ValueError("tol must be a float or 'auto'")
    handle = get_handle(N, 1, Wc.device.index if _is_torch(Wc) else None)
    res, _ = handle.isomp(Wc, dt, steps, tol=-1.0 if auto else float(tol), maxit=maxit, minit=minit,
                          compsum=bool(compsum), reinitialize=bool(reinitialize))
    st = res[0]
    if not inplace:   # non-contiguous input: copy the result back into the caller's array
        if _is_torch(W):
            W.copy_(Wc)
        else:
            W[...] = Wc
    if auto:
        if verbatim:
            print("Tolerance set to {}.".format(st['tol_used']))                     # :449-450
        if stats:
            stats['tol_auto'] = st['tol_used']                                       # :451-452
    if verbatim and steps > 0:
        print("Average number of iterations per step: {:.2f}".format(st['total_iterations'] / steps))   # :607-608
    if stats and steps > 0:                                                          # :609-611
        stats["iterations"] = st['total_iterations'] / steps
        stats["number_of_maxit"] = st['number_of_maxit'] / steps
    return W


# Default isospectral method (isospectral.py:617)
isomp = isomp_fixedpoint


def isomp_ensemble(W, dt, steps=100, stats=None, tol='auto', maxit=10, minit=1, compsum=False, reinitialize=False,
                   return_iterations=False):
    """Advance ``k`` INDEPENDENT simulations W[(k, N, N)] in one batched device call; every member has its own
    tolerance, convergence test and iteration counts — equivalent to ``k`` separate ``isomp`` calls
    (BASELINE config 5).  This is new functionality: the reference's (k, N, N) mode advects all members with
    member 0's stream function (quflow/laplacian/cpu.py:672-674).

    ``stats``: optional list that receives one dict per member with the reference's keys.
    """
    assert minit >= 1, "minit must be at least 1."
    assert maxit >= minit, "maxit must be at minit."
    if W.ndim != 3:
        raise ValueError("isomp_ensemble expects a (k, N, N) array")
    Wc = _prepare(W)
    k, N = Wc.shape[0], Wc.shape[-1]
    if k == 0:              # an empty member slice (more ranks than members) is a no-op
        if stats is not None:
            del stats[:]
        return (W, np.zeros((0, steps), dtype=np.int32)) if return_iterations else W
    auto = (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0)
    handle = get_handle(N, k, Wc.device.index if _is_torch(Wc) else None)
    res, iters = handle.isomp(Wc, dt, steps, tol=-1.0 if auto else float(tol), maxit=maxit, minit=minit,
                              compsum=bool(compsum), reinitialize=bool(reinitialize), want_iters=return_iterations)
    if Wc is not W:
        if _is_torch(W):
            W.copy_(Wc)
        else:
            W[...] = Wc
    if stats is not None:
        del stats[:]
        for st in res:
            d = {"iterations": st['total_iterations'] / max(steps, 1), "number_of_maxit": st['number_of_maxit'] / max(steps, 1)}
            if auto:
                d["tol_auto"] = st['tol_used']
            stats.append(d)
    if return_iterations:
        return W, iters
    return W
