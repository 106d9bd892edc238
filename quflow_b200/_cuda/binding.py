"""ctypes binding of include/quflow_b200.h.

This is the stub a quflow maintainer would add as ``quflow/_cuda/__init__.py``:
plain pointers and sizes across the boundary; torch tensors are used on the Python side only to own
device buffers and to provide the current CUDA stream.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBNAME = "libquflow_b200.so"

QF_OK = 0
QF_ERR_INVALID = -1
QF_ERR_CUDA = -2
QF_ERR_NONFINITE = -3
QF_ERR_NCCL = -4
QF_ERR_UNSUPPORTED = -5
QF_ERR_COMM = -6
QF_FLAG_COMPSUM = 1
QF_FLAG_REINITIALIZE = 2
QF_FLAG_MULTISTATE = 4
QF_FLAG_HOST_ROWS_OWN = 8
QF_UNIQUE_ID_BYTES = 128
QF_P2P_BLOB_BYTES = 256
QF_BUF_WHALF, QF_BUF_P, QF_BUF_SCRATCH = 0, 1, 2


class QfError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"quflow_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg


class qf_stats(ctypes.Structure):
    _fields_ = [("tol_used", ctypes.c_double), ("last_resnorm", ctypes.c_double),
                ("total_iterations", ctypes.c_int64), ("number_of_maxit", ctypes.c_int64),
                ("nonfinite", ctypes.c_int32), ("steps_done", ctypes.c_int32)]


class qf_phase_times(ctypes.Structure):
    _fields_ = [("poisson_ms", ctypes.c_float), ("gemm1_ms", ctypes.c_float), ("gemm2_ms", ctypes.c_float),
                ("post_ms", ctypes.c_float), ("update_ms", ctypes.c_float), ("x_tail_ms", ctypes.c_float),
                ("x_push_ms", ctypes.c_float), ("x_wait_ms", ctypes.c_float), ("x_mirror_ms", ctypes.c_float),
                ("x_control_ms", ctypes.c_float)]


_lib = None
_lock = threading.Lock()

# every symbol include/quflow_b200.h declares: (restype, argtypes)
_vp, _i, _d, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint
SYMBOLS = {
    "qf_device_count": (_i, []),
    "qf_version": (ctypes.c_char_p, []),
    "qf_last_error": (ctypes.c_char_p, []),
    "qf_create": (_i, [_i, _i, _i, ctypes.POINTER(_vp)]),
    "qf_destroy": (_i, [_vp]),
    "qf_solve_poisson": (_i, [_vp, _vp, _vp, _vp]),
    "qf_poisson_plan": (_i, [_i, ctypes.POINTER(_i), ctypes.POINTER(_i), _i]),
    "qf_laplace": (_i, [_vp, _vp, _vp, _vp]),
    "qf_norm_inf": (_i, [_vp, _vp, ctypes.POINTER(_d), _vp]),
    "qf_inner": (_i, [_vp, _vp, _vp, ctypes.POINTER(_d), _vp]),
    "qf_zgemm": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "qf_isomp": (_i, [_vp, _vp, _d, _i, _d, _i, _i, _u, ctypes.POINTER(qf_stats), ctypes.POINTER(ctypes.c_int32), _vp]),
    "qf_isomp_host": (_i, [_vp, _vp, _d, _i, _d, _i, _i, _u, ctypes.POINTER(qf_stats), ctypes.POINTER(ctypes.c_int32)]),
    "qf_solve_poisson_host": (_i, [_vp, _vp, _vp]),
    "qf_laplace_host": (_i, [_vp, _vp, _vp]),
    "qf_step_open": (_i, [_vp, _vp, _d, _d, _u, ctypes.POINTER(_d), _vp]),
    "qf_step_begin": (_i, [_vp, _vp, _vp]),
    "qf_step_buffer": (_vp, [_vp, _i]),
    "qf_step_hamiltonian": (_i, [_vp, _vp]),
    "qf_step_scale_p": (_i, [_vp, _i, _vp]),
    "qf_step_products": (_i, [_vp, _vp]),
    "qf_step_close_iteration": (_i, [_vp, _vp, _vp, _d, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_d), _vp]),
    "qf_step_increment": (_i, [_vp, _vp, _vp]),
    "qf_step_update": (_i, [_vp, _vp, _vp, _d, _vp]),
    "qf_step_stats": (_i, [_vp, ctypes.POINTER(qf_stats), _vp]),
    "qf_launch_count": (ctypes.c_int64, [_vp]),
    "qf_gemm_executed_flops": (_d, [_vp, _i]),
    "qf_gemm_is_3m": (_i, [_vp]),
    "qf_measure_fp64_tensor_peak": (_i, [_i, _i, ctypes.POINTER(_d), _vp]),
    "qf_profile_iteration": (_i, [_vp, _vp, _d, _i, ctypes.POINTER(qf_phase_times), _vp]),
    "qf_comm_get_unique_id": (_i, [_vp]),
    "qf_comm_init": (_i, [_vp, _vp, _i, _i]),
    "qf_set_emulated_ranks": (_i, [_vp, _i]),
    "qf_comm_p2p_export": (_i, [_vp, _vp]),
    "qf_comm_p2p_import": (_i, [_vp, _vp, _i, _i]),
    "qf_comm_set_tile": (_i, [_vp, _i]),
    "qf_comm_attach_local": (_i, [ctypes.POINTER(_vp), _i]),
    "qf_isomp_lockstep": (_i, [ctypes.POINTER(_vp), _i, ctypes.POINTER(_vp), _d, _i, _d, _i, _i, _u, ctypes.POINTER(qf_stats),
                               ctypes.POINTER(ctypes.c_int32), _vp]),
    "qf_set_fuse_post": (_i, [_vp, _i]),
    "qf_basis_size": (ctypes.c_longlong, [_i]),
    "qf_mat2shr": (_i, [_vp, _vp, _vp, _vp, ctypes.c_longlong, _vp]),
    "qf_shr2mat": (_i, [_vp, _vp, ctypes.c_longlong, _vp, _vp, _vp]),
    "qf_comm_mode": (_i, [_vp]),
}


def library_path() -> str:
    # QF_LIBRARY: developer override (instrumented builds made by tools/microbench); still a CUDA build, never a fallback
    return os.environ.get("QF_LIBRARY") or os.path.join(_HERE, _LIBNAME)


def library():
    """Load libquflow_b200.so (no CUDA call is made by loading).  Fails loudly if it was not built."""
    global _lib
    with _lock:
        if _lib is None:
            path = library_path()
            if not os.path.exists(path):
                raise ImportError(
                    f"{path} not found: the CUDA extension is not built. Run `python -c \"import __graft_entry__ as g; "
                    f"g.build()\"` (or `python quflow_b200/_cuda/build.py`). quflow_b200 has no CPU fallback.")
            lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
            for name, (res, args) in SYMBOLS.items():
                fn = getattr(lib, name)
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def _check(rc):
    if rc != QF_OK:
        raise QfError(rc, library().qf_last_error().decode())


def poisson_plan(N: int):
    """Work plan of the Poisson kernel (host code only): (params dict, units array (n, 8)), or None when N is too large
    for the band kernel (qf_create rejects such N)."""
    lib = library()
    params = (ctypes.c_int * 6)()
    n = lib.qf_poisson_plan(int(N), params, None, 0)
    if n < 0:
        raise QfError(n, lib.qf_last_error().decode())
    if n == 0:
        return None
    units = (ctypes.c_int * (8 * n))()
    lib.qf_poisson_plan(int(N), params, units, 8 * n)
    keys = ("L", "M", "NT", "CL", "PC", "nunits")
    return dict(zip(keys, params[:])), np.array(units[:], dtype=np.int32).reshape(n, 8)


def measure_fp64_tensor_peak(device: int = 0, reps: int = 5) -> float:
    """FP64 DMMA issue peak of `device` in TFLOP/s, measured now (roofline denominator of the GEMMs)."""
    out = _d(0.0)
    _check(library().qf_measure_fp64_tensor_peak(int(device), int(reps), ctypes.byref(out), _stream_ptr()))
    return out.value


def device_count() -> int:
    n = library().qf_device_count()
    return max(n, 0)


def _dev_ptr(t):
    """Device pointer of a contiguous complex128 torch CUDA tensor."""
    import torch
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.complex128 and t.is_contiguous()):
        raise TypeError("expected a contiguous complex128 CUDA tensor")
    return ctypes.c_void_p(t.data_ptr())


def _stream_ptr(device=None):
    """torch's current stream ON THE GIVEN DEVICE (default: the current device) as a cudaStream_t."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Handle:
    """Owns one qf_handle_t: scratch matrices, Laplacian factors, graphs for (N, batch, device)."""

    def __init__(self, N: int, batch: int = 1, device: int = 0):
        lib = library()
        if lib.qf_device_count() <= 0:
            raise QfError(QF_ERR_CUDA, "no CUDA device available (quflow_b200 has no CPU fallback)")
        self.N, self.batch, self.device = int(N), int(batch), int(device)
        h = ctypes.c_void_p()
        _check(lib.qf_create(self.N, self.batch, self.device, ctypes.byref(h)))
        self._h = h
        self._lib = lib
        self._busy = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.qf_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------
    def _stream(self):
        return _stream_ptr(self.device)     # the handle's device, whatever torch's current device is

    def use(self):
        """Context manager that pins the handle while a multi-call sequence (host-stepped driver, views of its buffers)
        is in flight: get_handle's LRU eviction never destroys a pinned handle."""
        handle = self

        class _Pin:
            def __enter__(self):
                handle._busy += 1
                return handle

            def __exit__(self, *exc):
                handle._busy -= 1
                return False
        return _Pin()

    def _shape_ok(self, a):
        shp = tuple(a.shape)
        want2, want3 = (self.N, self.N), (self.batch, self.N, self.N)
        if not (shp == want3 or (self.batch == 1 and shp == want2)):
            raise ValueError(f"array of shape {shp} does not match handle (batch={self.batch}, N={self.N})")

    @staticmethod
    def _host_ptr(a):
        if not (isinstance(a, np.ndarray) and a.dtype == np.complex128 and a.flags.c_contiguous):
            raise TypeError("expected a C-contiguous complex128 numpy array")
        return a.ctypes.data_as(ctypes.c_void_p)

    def launch_count(self) -> int:
        return int(self._lib.qf_launch_count(self._h))

    def gemm_info(self):
        """(3M arithmetic active?, executed flops of the full GEMM, executed flops of the upper-only GEMM)."""
        return (bool(self._lib.qf_gemm_is_3m(self._h)), float(self._lib.qf_gemm_executed_flops(self._h, 0)),
                float(self._lib.qf_gemm_executed_flops(self._h, 1)))

    # -- operators ---------------------------------------------------------------------
    def solve_poisson(self, W, out=None):
        self._shape_ok(W)
        if isinstance(W, np.ndarray):
            out = np.empty_like(W) if out is None else out
            _check(self._lib.qf_solve_poisson_host(self._h, self._host_ptr(W), self._host_ptr(out)))
            return out
        import torch
        out = torch.empty_like(W) if out is None else out
        _check(self._lib.qf_solve_poisson(self._h, _dev_ptr(W), _dev_ptr(out), self._stream()))
        return out

    def laplace(self, P, out=None):
        self._shape_ok(P)
        if isinstance(P, np.ndarray):
            out = np.empty_like(P) if out is None else out
            _check(self._lib.qf_laplace_host(self._h, self._host_ptr(P), self._host_ptr(out)))
            return out
        import torch
        out = torch.empty_like(P) if out is None else out
        _check(self._lib.qf_laplace(self._h, _dev_ptr(P), _dev_ptr(out), self._stream()))
        return out

    def norm_inf(self, W):
        self._shape_ok(W)
        out = (ctypes.c_double * self.batch)()
        _check(self._lib.qf_norm_inf(self._h, _dev_ptr(W), out, self._stream()))
        return np.array(out[:])

    def inner(self, P, W):
        """sum_ij Re(P_ij conj(W_ij)) per member (device tensors)."""
        self._shape_ok(P)
        self._shape_ok(W)
        out = (ctypes.c_double * self.batch)()
        _check(self._lib.qf_inner(self._h, _dev_ptr(P), _dev_ptr(W), out, self._stream()))
        return np.array(out[:])

    def zgemm(self, A, B, out=None):
        import torch
        self._shape_ok(A)
        self._shape_ok(B)
        out = torch.empty_like(A) if out is None else out
        _check(self._lib.qf_zgemm(self._h, _dev_ptr(A), _dev_ptr(B), _dev_ptr(out), self._stream()))
        return out

    def isomp(self, W, dt, steps, tol=-1.0, maxit=10, minit=1, compsum=False, reinitialize=False, want_iters=False,
              multistate=False, host_rows="all"):
        """Advance W in place.  Returns (list of per-member stats dicts, iters array or None).

        ``host_rows="own"`` (numpy W on a row-sharded handle, tile-exchange path): the host array is a row-distributed
        state — this rank reads and writes its own two row blocks only (``quflow_b200.distributed.row_blocks``), the
        other rows of the array are left untouched.

        Raises ValueError on a non-finite residual (the reference raises it from scipy.linalg.norm).
        """
        self._shape_ok(W)
        flags = ((QF_FLAG_COMPSUM if compsum else 0) | (QF_FLAG_REINITIALIZE if reinitialize else 0)
                 | (QF_FLAG_MULTISTATE if multistate else 0))
        if host_rows == "own":
            if not isinstance(W, np.ndarray):
                raise TypeError("host_rows='own' applies to numpy (host) arrays")
            flags |= QF_FLAG_HOST_ROWS_OWN
        elif host_rows != "all":
            raise ValueError("host_rows must be 'all' or 'own'")
        stats = (qf_stats * self.batch)()
        iters = np.zeros((self.batch, max(steps, 1)), dtype=np.int32) if want_iters else None
        iters_p = iters.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)) if want_iters else None
        if isinstance(W, np.ndarray):
            rc = self._lib.qf_isomp_host(self._h, self._host_ptr(W), float(dt), int(steps), float(tol), int(maxit),
                                         int(minit), flags, stats, iters_p)
        else:
            rc = self._lib.qf_isomp(self._h, _dev_ptr(W), float(dt), int(steps), float(tol), int(maxit), int(minit),
                                    flags, stats, iters_p, self._stream())
        if rc == QF_ERR_NONFINITE:
            raise ValueError("array must not contain infs or NaNs")
        if rc == QF_ERR_INVALID:
            raise AssertionError(self._lib.qf_last_error().decode())
        _check(rc)
        out = [dict(tol_used=s.tol_used, last_resnorm=s.last_resnorm, total_iterations=int(s.total_iterations),
                    number_of_maxit=int(s.number_of_maxit), steps_done=int(s.steps_done)) for s in stats]
        if want_iters:
            iters = iters[:, :steps]
        return out, iters

    # -- host-stepped driver (include/quflow_b200.h, "host-stepped driver") -----------------
    def buffer_tensor(self, which: int):
        """A torch view (N, N) complex128 of one of the handle's device buffers (QF_BUF_*)."""
        import torch
        ptr = self._lib.qf_step_buffer(self._h, int(which))
        if not ptr:
            raise QfError(QF_ERR_INVALID, f"qf_step_buffer({which}) returned NULL")

        class _View:
            __cuda_array_interface__ = {"shape": (self.N, self.N), "typestr": "<c16", "data": (int(ptr), False),
                                        "version": 2, "strides": None}
        return torch.as_tensor(_View(), device=f"cuda:{self.device}")

    def step_open(self, W, dt, tol=-1.0, compsum=False, reinitialize=False) -> float:
        flags = (QF_FLAG_COMPSUM if compsum else 0) | (QF_FLAG_REINITIALIZE if reinitialize else 0)
        used = _d(0.0)
        _check(self._lib.qf_step_open(self._h, _dev_ptr(W), float(dt), float(tol), flags, ctypes.byref(used), self._stream()))
        return used.value

    def step_begin(self, W):
        _check(self._lib.qf_step_begin(self._h, _dev_ptr(W), self._stream()))

    def step_hamiltonian(self):
        _check(self._lib.qf_step_hamiltonian(self._h, self._stream()))

    def step_scale_p(self, divide: bool):
        _check(self._lib.qf_step_scale_p(self._h, 1 if divide else 0, self._stream()))

    def step_products(self):
        _check(self._lib.qf_step_products(self._h, self._stream()))

    def step_close_iteration(self, W, F, fscale, maxit, minit):
        """Returns (loop continues?, residual).  Raises ValueError on a non-finite residual like scipy.linalg.norm."""
        active, res = _i(0), _d(0.0)
        rc = self._lib.qf_step_close_iteration(self._h, _dev_ptr(W), _dev_ptr(F) if F is not None else None, float(fscale),
                                               int(maxit), int(minit), ctypes.byref(active), ctypes.byref(res), self._stream())
        if rc == QF_ERR_NONFINITE:
            raise ValueError("array must not contain infs or NaNs")
        _check(rc)
        return bool(active.value), res.value

    def step_increment(self, out):
        _check(self._lib.qf_step_increment(self._h, _dev_ptr(out), self._stream()))
        return out

    def step_update(self, W, F, fscale):
        rc = self._lib.qf_step_update(self._h, _dev_ptr(W), _dev_ptr(F) if F is not None else None, float(fscale), self._stream())
        if rc == QF_ERR_UNSUPPORTED:
            raise NotImplementedError(self._lib.qf_last_error().decode())
        _check(rc)

    def step_stats(self):
        st = qf_stats()
        _check(self._lib.qf_step_stats(self._h, ctypes.byref(st), self._stream()))
        return dict(tol_used=st.tol_used, last_resnorm=st.last_resnorm, total_iterations=int(st.total_iterations),
                    number_of_maxit=int(st.number_of_maxit), steps_done=int(st.steps_done))

    def profile_iteration(self, W, dt, reps=5):
        pt = qf_phase_times()
        _check(self._lib.qf_profile_iteration(self._h, _dev_ptr(W), float(dt), int(reps), ctypes.byref(pt), self._stream()))
        return {k: getattr(pt, k) for k, _ in qf_phase_times._fields_}

    # -- multi-GPU -----------------------------------------------------------------------
    def set_emulated_ranks(self, nranks: int):
        _check(self._lib.qf_set_emulated_ranks(self._h, int(nranks)))

    def p2p_export(self) -> bytes:
        buf = ctypes.create_string_buffer(QF_P2P_BLOB_BYTES)
        _check(self._lib.qf_comm_p2p_export(self._h, buf))
        return buf.raw

    def p2p_import(self, blobs, rank: int, nranks: int):
        raw = b"".join(blobs)
        assert len(raw) == nranks * QF_P2P_BLOB_BYTES
        _check(self._lib.qf_comm_p2p_import(self._h, ctypes.create_string_buffer(raw, len(raw)), int(rank), int(nranks)))

    def comm_mode(self) -> str:
        return {0: "none", 1: "nccl", 2: "pull", 5: "tile"}[int(self._lib.qf_comm_mode(self._h))]

    def comm_set_tile(self, enable: bool):
        """True: tile exchange (sharded tail and update, W~ pushed by the owners); False: pull all-gather of A and S."""
        _check(self._lib.qf_comm_set_tile(self._h, 1 if enable else 0))

    def set_fuse_post(self, enable: bool):
        """Run the tail of the iteration fused into the epilogue of the second GEMM (True) or as its own kernel (False)."""
        _check(self._lib.qf_set_fuse_post(self._h, 1 if enable else 0))

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = ctypes.create_string_buffer(unique_id, QF_UNIQUE_ID_BYTES)
        _check(self._lib.qf_comm_init(self._h, buf, int(rank), int(nranks)))


def attach_local(handles):
    """Test hook: attach handles of this process (same device, same N) to each other as the ranks of one tile-exchange
    group; drive them with :func:`isomp_lockstep`."""
    arr = (_vp * len(handles))(*[h._h for h in handles])
    _check(library().qf_comm_attach_local(arr, len(handles)))


def isomp_lockstep(handles, Ws, dt, steps, tol=-1.0, maxit=10, minit=1, compsum=False, reinitialize=False):
    """Advance the replicated states ``Ws`` (one torch CUDA tensor per rank) of a local tile-exchange group in lock step
    on one GPU.  Returns (list of stats dicts, iteration counts (G, steps))."""
    G = len(handles)
    flags = (QF_FLAG_COMPSUM if compsum else 0) | (QF_FLAG_REINITIALIZE if reinitialize else 0)
    hs = (_vp * G)(*[h._h for h in handles])
    ws = (_vp * G)(*[_dev_ptr(W) for W in Ws])
    stats = (qf_stats * G)()
    iters = np.zeros((G, max(steps, 1)), dtype=np.int32)
    rc = library().qf_isomp_lockstep(hs, G, ws, float(dt), int(steps), float(tol), int(maxit), int(minit), flags, stats,
                                     iters.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), _stream_ptr(handles[0].device))
    if rc == QF_ERR_NONFINITE:
        raise ValueError("array must not contain infs or NaNs")
    _check(rc)
    out = [dict(tol_used=s.tol_used, last_resnorm=s.last_resnorm, total_iterations=int(s.total_iterations),
                number_of_maxit=int(s.number_of_maxit), steps_done=int(s.steps_done)) for s in stats]
    return out, iters[:, :steps]


def comm_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(QF_UNIQUE_ID_BYTES)
    _check(library().qf_comm_get_unique_id(buf))
    return buf.raw


from collections import OrderedDict  # noqa: E402

_handles: "OrderedDict[tuple, Handle]" = OrderedDict()
_MAX_CACHED_HANDLES = 8      # a handle owns ~8 N^2 complex matrices (0.55 GB at N=2048): keep the cache bounded


def get_handle(N: int, batch: int = 1, device: int | None = None) -> Handle:
    """Module-level LRU cache, mirroring the reference's per-N caches (quflow/laplacian/cpu.py:11-17).

    Callers use the handle for the duration of one call (or pin it with ``Handle.use()`` across several); the least
    recently used idle one is destroyed when more than ``_MAX_CACHED_HANDLES`` distinct (N, batch, device) combinations
    have been seen.  One handle serves one host thread / stream at a time.  Hold your own ``Handle`` for long-lived use
    (bench.py, multi-GPU sharding)."""
    if device is None:
        try:
            import torch
            device = torch.cuda.current_device() if torch.cuda.is_available() else 0
        except Exception:
            device = 0
    key = (int(N), int(batch), int(device))
    if key in _handles:
        _handles.move_to_end(key)
        return _handles[key]
    if len(_handles) >= _MAX_CACHED_HANDLES:
        for old_key in [k for k, hd in _handles.items() if hd._busy == 0][:len(_handles) - _MAX_CACHED_HANDLES + 1]:
            _handles.pop(old_key).close()      # least recently used first; handles pinned by Handle.use() are kept
    _handles[key] = Handle(*key)
    return _handles[key]
