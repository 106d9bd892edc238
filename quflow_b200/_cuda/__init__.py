"""quflow._cuda equivalent: thin ctypes binding to the sm_100a library (libquflow_b200.so).

There is deliberately no CPU fallback: importing this package works anywhere, but any call
raises unless the compiled library is present and a CUDA device is available.
"""
from .binding import (  # noqa: F401
    QfError, Handle, get_handle, library, library_path, device_count, QF_FLAG_COMPSUM, QF_FLAG_REINITIALIZE,
)
