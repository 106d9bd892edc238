"""Matrix <-> real spherical-harmonic coefficients on the B200 — the data formats on either side of the hot path.

Reference: quflow/quantization.py — ``mat2shr`` (:488-519, kernel ``mat2shr_parallel_`` :283-325) and ``shr2mat``
(:440-485, kernel ``shr2mat_parallel_`` :172-227).  Same names, argument meaning and coefficient ordering
(``omega[el**2 + el + m]``, quflow/utils.py:91-105); only the ``'mat'``/``'shr'`` pair is covered (``fun``/``shc`` need the
spherical-harmonic transforms of the reference).

The quantization basis is NOT computed here: pass the reference's own array (``quflow.quantization.get_basis(N)``, an
eigenproblem per m that the reference caches on disk) — as a numpy array, or once as a CUDA tensor (``device_basis``) so
that spectral output of a device-resident state never touches the host matrix.
"""
import ctypes

import numpy as np

from ._cuda import get_handle
from ._cuda.binding import library, _check, _stream_ptr

__all__ = ["mat2shr", "shr2mat", "device_basis", "basis_size"]


def basis_size(N: int) -> int:
    """Number of doubles of the basis for size N: sum over m of (N-m)**2 (quantization.py:25-42)."""
    return int(library().qf_basis_size(int(N)))


def device_basis(basis, device=None):
    """Upload the reference's flat basis array once; returns a float64 CUDA tensor to pass as ``basis=``."""
    import torch
    if isinstance(basis, torch.Tensor):
        return basis.to(device="cuda" if device is None else device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(basis, dtype=np.float64)).to("cuda" if device is None else device)


def _dev(t):
    return ctypes.c_void_p(t.data_ptr())


def mat2shr(W, basis, elmax=-1):
    """Real spherical-harmonic coefficients of the (N, N) complex matrix ``W`` (reference: quantization.py:488-519).

    ``W``: numpy array or CUDA tensor; the result is of the same kind, float64, of length ``N**2`` or
    ``(elmax+1)**2`` when ``elmax > 0``."""
    import torch
    on_host = isinstance(W, np.ndarray)
    Wd = torch.from_numpy(np.ascontiguousarray(W, dtype=np.complex128)).cuda() if on_host else W.contiguous()
    if Wd.dtype != torch.complex128 or Wd.ndim != 2 or Wd.shape[0] != Wd.shape[1]:
        raise TypeError("W must be an (N, N) complex128 array")
    N = Wd.shape[-1]
    B = device_basis(basis, Wd.device)
    if B.numel() != basis_size(N):
        raise ValueError(f"basis has {B.numel()} entries, expected {basis_size(N)} for N={N}")
    n = N * N if elmax <= 0 else min(elmax + 1, N) ** 2
    omega = torch.empty(n, dtype=torch.float64, device=Wd.device)
    h = get_handle(N, 1, Wd.device.index)
    _check(library().qf_mat2shr(h._h, _dev(Wd), _dev(B), _dev(omega), ctypes.c_longlong(n), _stream_ptr(Wd.device)))
    return omega.cpu().numpy() if on_host else omega


def shr2mat(omega, basis, N=-1):
    """Matrix of the real spherical-harmonic coefficients ``omega`` (reference: quantization.py:440-485).

    ``omega``: float64 numpy array or CUDA tensor of length ``(elmax+1)**2``; ``N = -1`` means ``N = elmax + 1``."""
    import torch
    on_host = isinstance(omega, np.ndarray)
    od = torch.from_numpy(np.ascontiguousarray(omega, dtype=np.float64)).cuda() if on_host else omega.contiguous()
    if od.dtype != torch.float64 or od.ndim != 1:
        raise TypeError("omega must be a 1-D float64 array")
    if N == -1:
        N = int(round(np.sqrt(od.numel())))
    B = device_basis(basis, od.device)
    if B.numel() != basis_size(N):
        raise ValueError(f"basis has {B.numel()} entries, expected {basis_size(N)} for N={N}")
    W = torch.empty((N, N), dtype=torch.complex128, device=od.device)
    h = get_handle(N, 1, od.device.index)
    _check(library().qf_shr2mat(h._h, _dev(od), ctypes.c_longlong(od.numel()), _dev(B), _dev(W), _stream_ptr(od.device)))
    return W.cpu().numpy() if on_host else W
