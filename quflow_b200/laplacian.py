"""Hoppe–Yau Laplacian operators on the B200 — drop-in for quflow.laplacian (dense skew-Hermitian path).

Reference: quflow/laplacian/cpu.py — ``solve_poisson`` (:681-734), ``laplace`` (:628-669),
``laplacian`` (:604-625), ``select_first`` (:672-674).  All arithmetic runs in the CUDA library
(`quflow_b200/csrc/poisson.cu`); there is no CPU path.
"""
import numpy as np

from ._cuda import get_handle


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def select_first(W):
    """Reduce a multi-state (k, N, N) array to its first member — cpu.py:672-674."""
    zeroind = (0,) * (W.ndim - 2) + (Ellipsis,)
    if _is_torch(W):
        return W[zeroind].contiguous()
    return np.ascontiguousarray(W[zeroind])


def _prepare(W):
    if _is_torch(W):
        import torch
        if W.dtype != torch.complex128:
            raise TypeError("quflow_b200 computes in complex128 only")
        return W.contiguous()
    W = np.asarray(W)
    if W.dtype != np.complex128:
        raise TypeError("quflow_b200 computes in complex128 only (got %s)" % W.dtype)
    return np.ascontiguousarray(W)


def solve_poisson(W, reduce=select_first):
    """Solve ΔP = W for the stream matrix P (cpu.py:681-734).

    ``W``: (N, N) or (k, N, N) complex128 numpy array (host; copied to the GPU and back) or torch
    CUDA tensor (stays on the device).  Like the reference, only the upper triangle of ``W`` is read,
    ``tr(W)/N`` is removed, and ``P`` comes back trace-free and exactly skew-Hermitian.  Unlike the
    reference a fresh array is returned (the reference returns a module-level cache, cpu.py:726).

    The signature has no ``time`` parameter on purpose: ``isomp`` probes ``hamiltonian(W, time=...)`` and
    treats the TypeError as "autonomous" (isospectral.py:416-423).
    """
    if W.ndim >= 3:
        W = reduce(W)
    W = _prepare(W)
    N = W.shape[-1]
    return get_handle(N, 1, W.device.index if _is_torch(W) else None).solve_poisson(W)


def laplace(P):
    """Apply the quantised Laplacian, W = ΔP (cpu.py:628-669, dense branch)."""
    P = _prepare(P)
    N = P.shape[-1]
    return get_handle(N, 1, P.device.index if _is_torch(P) else None).laplace(P)


def laplacian(N, bc=False, dtype=np.float64):
    """The (N, N, 2) tridiagonal coefficient table (cpu.py:55-95, 604-625) — host-side convenience only;
    the device keeps its own factorised copy."""
    i, j = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    m = np.abs(j - i).astype(dtype)
    k = np.minimum(i, j).astype(dtype)
    lap = np.zeros((N, N, 2), dtype=dtype)
    lap[..., 0] = -((N - 1) * (2 * k + 1 + m) - 2 * k * (k + m))
    lap[..., 1] = np.sqrt(((k + m) * (N - k - m)) * (k * (N - k)))
    if bc:
        lap[0, 0, 0] -= 0.5
    return lap
