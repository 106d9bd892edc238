"""Loggers and Sobolev-type inner products on the device — drop-in for the functions of quflow/physics.py that sit on
either side of the hot path (they are what `qf.solve` callbacks evaluate at every output step).

Reference: quflow/physics.py:9-38.  All of them accept numpy arrays (copied to the GPU) or torch CUDA tensors (no
copy); the Poisson solve / Laplacian and the reductions run in the CUDA library.
"""
import numpy as np

from .geometry import inner_L2
from .laplacian import solve_poisson, laplace, _is_torch


def _on_device(W):
    import torch
    if _is_torch(W):
        return W
    return torch.from_numpy(np.ascontiguousarray(W, dtype=np.complex128)).to("cuda")


def inner_Hm1(W1, W2):
    """H^-1 inner product -<W1, solve_poisson(W2)>_L2 — physics.py:9-11."""
    W2d = _on_device(W2)
    return -inner_L2(W1, solve_poisson(W2d))


def norm_Hm1(W):
    """physics.py:13-14."""
    return float(np.sqrt(inner_Hm1(W, W)))


def inner_H1(P1, P2):
    """H^1 inner product -<P1, laplace(P2)>_L2 — physics.py:16-18."""
    P2d = _on_device(P2)
    return -inner_L2(P1, laplace(P2d))


def norm_H1(P):
    """physics.py:20-21."""
    return float(np.sqrt(inner_H1(P, P)))


def energy_euler(W):
    """Energy of the 2-D Euler state with vorticity matrix W: -<W, solve_poisson(W)>_L2 / 2 — physics.py:26-32."""
    Wd = _on_device(W)
    return -inner_L2(Wd, solve_poisson(Wd)) / 2.0


def enstrophy(W):
    """Enstrophy <W, W>_L2 / 2 — physics.py:34-38."""
    return inner_L2(W, W) / 2.0
