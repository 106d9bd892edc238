"""Isospectral midpoint integrator on the B200 — drop-in for ``quflow.integrators.isomp``.

Reference: quflow/integrators/isospectral.py:338-613 (``isomp_fixedpoint``; alias ``isomp`` :617).
The signature, argument meaning, in-place semantics, ``stats`` keys and error behaviour mirror the
reference.  The whole fixed-point loop runs inside the CUDA library (csrc/isomp.cu) without host
synchronisation; Python only validates arguments and converts the statistics.

Two execution modes share the same CUDA kernels:

* the fused mode (default arguments): the whole run is enqueued by one ``qf_isomp`` call; the GPU decides the
  iteration counts and the host synchronises once per call;
* the host-stepped mode, taken when the caller passes hooks that run host code inside the step — ``callback``,
  ``forcing``, ``strang_splitting``, a custom or time-dependent ``hamiltonian``: the same kernels are driven one
  fixed-point iteration at a time through the ``qf_step_*`` entry points, the hooks receive (and return) arrays of
  the same kind as ``W`` (numpy on the host, torch tensors on the device).
"""
import numpy as np

from ._cuda import get_handle
from ._cuda.binding import QF_BUF_WHALF, QF_BUF_P, QF_BUF_SCRATCH
from .geometry import hbar
from .laplacian import solve_poisson, _is_torch, _prepare


def _is_default_hamiltonian(hamiltonian):
    """The default Hamiltonian is recognised by identity, like the reference's own GPU seam
    (quflow/simulation.py:554); it runs fused on the device."""
    if hamiltonian is None or hamiltonian is solve_poisson:
        return True
    mod = getattr(hamiltonian, "__module__", "") or ""
    if getattr(hamiltonian, "__name__", "") == "solve_poisson" and mod.startswith("quflow.laplacian"):
        return True   # the reference's own default, when this package is plugged in behind quflow
    return False


def _isomp_host_stepped(W, dt, steps, hamiltonian, time, forcing, strang_splitting, stats, callback, tol, maxit,
                        minit, verbatim, compsum, reinitialize, handle=None):
    """isomp_fixedpoint with host code inside the step (isospectral.py:403-424, 466-467, 488-491, 511-520, 550-551,
    594-603).  The matrices stay on the GPU; what the hooks see is of the same kind as the caller's ``W``.

    ``handle``: a row-sharded handle (``distributed.ShardedIsomp``): the two GEMMs of every iteration run sharded over the
    ranks, A and S are completed on every rank, and everything the hooks see is complete and identical on every rank."""
    import torch
    torch_in = _is_torch(W)
    if torch_in:
        Wd = _prepare(W)
        dev = Wd.device
    else:
        Wh_in = _prepare(W)
        dev = torch.device("cuda", torch.cuda.current_device())
        Wd = torch.from_numpy(Wh_in).to(dev)
    N = Wd.shape[-1]

    def to_user(t, writable=False):
        if torch_in:
            return t
        a = t.cpu().numpy()
        if not writable:
            a.flags.writeable = False    # a copy: writes would be lost, so make them fail loudly
        return a

    def from_user(x, what):
        if _is_torch(x):
            t = x.to(device=dev, dtype=torch.complex128)
        else:
            t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np.complex128)).to(dev)
        if tuple(t.shape) != (N, N):
            raise ValueError(f"{what} returned an array of shape {tuple(t.shape)}, expected {(N, N)}")
        return t.contiguous()

    default_ham = _is_default_hamiltonian(hamiltonian)
    # autonomy probes, exactly as the reference does them (isospectral.py:403-424)
    autonomous_force = True
    if forcing is not None and time is not None:
        try:
            forcing(to_user(Wd), to_user(Wd), time=time)
        except TypeError:
            pass
        else:
            autonomous_force = False
    autonomous = True
    if time is not None and not default_ham:       # the default solve_poisson has no `time` keyword
        try:
            hamiltonian(to_user(Wd), time=time)
        except TypeError:
            pass
        else:
            autonomous = False

    auto = (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0)   # :440
    if isinstance(tol, str) and not auto:
        raise ValueError("tol must be a float or 'auto'")
    h = handle if handle is not None else get_handle(N, 1, dev.index)
    with h.use():      # pinned: hooks may create other handles while views of this one's buffers are alive
        tol_used = h.step_open(Wd, dt, -1.0 if auto else float(tol), compsum=bool(compsum), reinitialize=bool(reinitialize))
        if auto:
            if verbatim:
                print("Tolerance set to {}.".format(tol_used))            # :449-450
            if stats:
                stats['tol_auto'] = tol_used                               # :451-452
        Whalf = h.buffer_tensor(QF_BUF_WHALF)
        Phalf = h.buffer_tensor(QF_BUF_P)
        inc = h.buffer_tensor(QF_BUF_SCRATCH)

        def strang(Wcur):
            Wnew = strang_splitting(dt / 2, to_user(Wcur, writable=True))   # :466-467, :602-603 (the reference rebinds W)
            if Wnew is not Wcur:
                Wcur.copy_(from_user(Wnew, "strang_splitting"))
            return Wcur

        seen_maxit = 0
        for k in range(steps):
            if strang_splitting:
                Wd = strang(Wd)
            h.step_begin(Wd)                                               # :470-472, :481-482
            FW = None
            for i in range(maxit):                                         # :475
                if default_ham:
                    h.step_hamiltonian()                                   # :489, :492
                else:
                    Pu = hamiltonian(to_user(Whalf)) if autonomous else hamiltonian(to_user(Whalf), time=time + dt / 2)   # :488-491
                    Phalf.copy_(from_user(Pu, "hamiltonian"))
                    h.step_scale_p(divide=False)                           # :492
                h.step_products()                                          # :496, :499
                if forcing:                                                # :511-520
                    h.step_scale_p(divide=True)                            # :513
                    Fu = (forcing(to_user(Phalf), to_user(Whalf)) if autonomous_force
                          else forcing(to_user(Phalf), to_user(Whalf), time=time + dt / 2))
                    FW = from_user(Fu, "forcing")
                active, _ = h.step_close_iteration(Wd, FW, dt / 2, maxit, minit)   # :503-509, :518-536
                if not active:
                    break
            if verbatim:
                n_maxit = h.step_stats()['number_of_maxit']
                if n_maxit > seen_maxit:
                    print("Max iterations {} reached at step {}.".format(maxit, k))      # :538-542
                seen_maxit = n_maxit
            if callback is not None:                                       # :547-551
                h.step_increment(inc)
                callback(to_user(Wd), to_user(inc))
            if compsum and forcing:
                raise NotImplementedError("Compensated sum with forcing is not yet implemented.")   # :588-589
            h.step_update(Wd, FW, dt / 2)                                  # :553-596
            if time is not None:
                time += dt                                                 # :598-599
            if strang_splitting:
                Wd = strang(Wd)

        st = h.step_stats()
    if verbatim and steps > 0:
        print("Average number of iterations per step: {:.2f}".format(st['total_iterations'] / steps))   # :607-608
    if stats and steps > 0:                                            # :609-611
        stats["iterations"] = st['total_iterations'] / steps
        stats["number_of_maxit"] = st['number_of_maxit'] / steps
    if torch_in:
        if Wd.data_ptr() != W.data_ptr():
            W.copy_(Wd)
    else:
        W[...] = Wd.cpu().numpy()
    return W


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

    ``W``: skew-Hermitian (N, N) complex128, or (k, N, N) for the reference's multi-state mode (members 1.. are
    advected passively by member 0's stream function, cpu.py:672-674); a numpy array is overwritten in place and returned, exactly like
    the reference (host→device and back once per call); a torch CUDA tensor is advanced in place on the device.
    ``time`` is accepted (``qf.solve`` always passes it, simulation.py:727) — the default Hamiltonian is
    autonomous so it has no effect on it.  ``callback(W, dW)``, ``forcing(P, W[, time])``,
    ``strang_splitting(dt/2, W)`` and custom / time-dependent ``hamiltonian(W[, time])`` are honoured with the
    reference's semantics in the host-stepped mode (module docstring); every matrix on the path is assumed
    skew-Hermitian, as the reference's own solve_poisson and conj_subtract_ assume.
    """
    assert minit >= 1, "minit must be at least 1."          # isospectral.py:400
    assert maxit >= minit, "maxit must be at minit."         # isospectral.py:401
    hooks = (forcing is not None or strang_splitting is not None or callback is not None
             or not _is_default_hamiltonian(hamiltonian))
    if W.ndim != 2 and (W.ndim != 3 or hooks):
        raise NotImplementedError("multi-state input must be (k, N, N) and runs with the default Hamiltonian and without "
                                  "hooks; use quflow_b200.isomp_ensemble for independent members")
    if hooks:
        return _isomp_host_stepped(W, dt, steps, hamiltonian, time, forcing, strang_splitting, stats, callback, tol,
                                   maxit, minit, verbatim, compsum, reinitialize)
    Wc = _prepare(W)
    inplace = Wc is W or (_is_torch(W) and Wc.data_ptr() == W.data_ptr())
    N = Wc.shape[-1]
    auto = (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0)   # :440
    if isinstance(tol, str) and not auto:
        raise ValueError("tol must be a float or 'auto'")
    # (k, N, N): ONE multi-state run — members 1.. are advected by member 0's stream function (select_first,
    # cpu.py:672-674); tolerance, residual and statistics come from member 0 (isospectral.py:444-446, 528-531)
    k = Wc.shape[0] if Wc.ndim == 3 else 1
    handle = get_handle(N, k, Wc.device.index if _is_torch(Wc) else None)
    res, _ = handle.isomp(Wc, dt, steps, tol=-1.0 if auto else float(tol), maxit=maxit, minit=minit,
                          compsum=bool(compsum), reinitialize=bool(reinitialize), multistate=Wc.ndim == 3)
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
