"""Per-CTA phase timeline of k_poisson_band (needs a -DQF_PTRACE build: see build_trace() below).

    python tools/microbench/poisson_trace.py [N]      # on a GPU box
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "quflow_b200", "_cuda", "libquflow_b200_trace.so")


def build_trace():
    csrc = os.path.join(ROOT, "quflow_b200", "csrc")
    srcs = [os.path.join(csrc, f) for f in ("api.cu", "poisson.cu", "zgemm.cu", "isomp.cu", "comm.cu")]
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17",
                           "-lineinfo", "-DQF_PTRACE", "-Xcompiler", "-fPIC", "-shared", "-o", TRACE_LIB] + srcs)
    return TRACE_LIB


if __name__ == "__main__":
    if "--build" in sys.argv:
        print(build_trace())
        sys.exit(0)
    os.environ["QF_LIBRARY"] = TRACE_LIB
    import numpy as np
    import torch
    from quflow_b200._cuda import binding
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    h = binding.get_handle(N)
    A = torch.randn(N, N, dtype=torch.complex128, device="cuda")
    W = (A - A.conj().T).contiguous()
    P = torch.empty_like(W)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        flush.zero_()
        h.solve_poisson(W, out=P)
    torch.cuda.synchronize()
    lib = binding.library()
    n = 4096 * 16
    buf = (ctypes.c_ulonglong * n)()
    lib.qf_ptrace_read.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.qf_ptrace_read(buf, n) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(4096, 16)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t[:, :14] - t0) / 1e3
    print(f"N={N} CTAs traced: {len(t)}  kernel span {rel[:, 13].max():.1f} us")
    names = ["start", "loads issued", "pass1+scan (loads landed)", "barrier", "cluster xchg F", "prefix+pass2", "barrier",
             "bwd pass1+scan", "barrier", "cluster xchg B", "prefix+bwd pass2", "direct st + STS", "barrier", "mirror st"]
    d = np.diff(rel, axis=1)
    print("phase durations (us): mean / p50 / max over CTAs, then for CTA 0, 1, 300, 600")
    for k in range(13):
        col = d[:, k]
        pick = [col[i] if i < len(col) else float('nan') for i in (0, 1, 300, 600)]
        print(f"  {names[k + 1]:28s} {col.mean():6.2f} {np.median(col):6.2f} {col.max():6.2f}   " + " ".join(f"{x:6.2f}" for x in pick))
    early = rel[:, 0] < 5.0
    print("mean phase durations, CTAs started before 5 us (%d) vs later (%d):" % (early.sum(), (~early).sum()))
    for k in range(13):
        print(f"  {names[k + 1]:28s} {d[early, k].mean():6.2f} {d[~early, k].mean() if (~early).any() else float('nan'):6.2f}")
    life = rel[:, 13] - rel[:, 0]
    print(f"CTA lifetime: mean {life.mean():.2f} p50 {np.median(life):.2f} max {life.max():.2f} us")
    order = np.argsort(rel[:, 0])
    print("start times (us) of every 64th CTA in start order:", np.round(rel[order[::64], 0], 1))
    sm = t[:, 15]
    for q in (0, 1, 2):
        ids = np.where(sm == q)[0]
        print(f"SM {q}: " + ", ".join(f"cta{idx}:[{rel[idx, 0]:.1f}-{rel[idx, 13]:.1f}]" for idx in ids))
