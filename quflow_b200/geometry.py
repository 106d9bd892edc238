"""Geometry helpers on the hot path (reference: quflow/geometry.py)."""
import numpy as np


def hbar(N):
    """Planck-like constant of the quantisation, 2/sqrt(N^2-1) — quflow/geometry.py:7-9."""
    return 2.0 / np.sqrt(float(N) ** 2 - 1.0)


def _device_pair(P, W):
    """Both operands as contiguous complex128 CUDA tensors on one device (numpy inputs are copied over)."""
    import torch
    from .laplacian import _is_torch, _prepare
    dev = None
    for x in (P, W):
        if _is_torch(x):
            dev = x.device
    if dev is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    out = []
    for x in (P, W):
        x = _prepare(x)
        out.append(x if _is_torch(x) else torch.from_numpy(x).to(dev))
    if out[0].shape != out[1].shape or out[0].ndim != 2:
        raise ValueError("inner_L2 expects two (N, N) matrices of the same shape")
    return out[0], out[1]


def inner_L2(P, W):
    """L2 inner product Re sum(P conj(W)) / N — quflow/geometry.py:72-76 (dense branch), computed on the device with a
    fixed summation tree (deterministic)."""
    from ._cuda import get_handle
    Pd, Wd = _device_pair(P, W)
    N = Wd.shape[-1]
    return float(get_handle(N, 1, Wd.device.index).inner(Pd, Wd)[0]) / N


def norm_L2(W):
    """Scaled Frobenius norm ||W||_F / sqrt(N) — quflow/geometry.py:53-68 (dense branch)."""
    return float(np.sqrt(inner_L2(W, W)))
