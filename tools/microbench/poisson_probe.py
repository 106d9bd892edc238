"""Poisson-solve probe: parity against the CPU oracle and kernel timing (run on a GPU box through gpurun).

    python tools/microbench/poisson_probe.py [--check] [--sizes 512,1024,2048]

Timing: `reps` launches per CUDA-event pair, rotating over enough (W, P) buffer pairs that every launch
finds its inputs evicted from the 126 MB L2 (cold, like inside the isomp loop where the GEMMs run in between),
and the same with one buffer pair (warm).  Reports us per launch and 32 N^2 B / time.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import quflow_b200 as qf          # noqa: E402
from quflow_b200._cuda import binding   # noqa: E402


def check(sizes):
    import oracle
    worst = 0.0
    for N in sizes:
        W = oracle.random_skewherm(N, seed=N) + (0.3j * np.eye(N) if N % 2 else 0)
        P = qf.solve_poisson(torch.from_numpy(W).cuda()).cpu().numpy()
        Pref = oracle.solve_poisson(W)
        err = np.linalg.norm(P - Pref) / np.linalg.norm(Pref)
        skew = np.abs(P + P.conj().T).max()
        worst = max(worst, err)
        flag = "" if (err < 1e-13 and skew == 0.0) else "   <-- FAIL"
        print(f"check N={N:5d} rel.err={err:.2e} skew={skew:.1e} tr={abs(np.trace(P)):.1e}{flag}", flush=True)
    return worst


def timeit(N, reps=20, rounds=5):
    h = binding.get_handle(N)
    nbuf = max(2, int(np.ceil(400e6 / (32.0 * N * N))))
    g = torch.Generator(device="cuda").manual_seed(1)
    Ws = []
    for _ in range(nbuf):
        A = torch.randn(N, N, dtype=torch.complex128, device="cuda", generator=g)
        Ws.append((A - A.conj().T).contiguous())
    Ps = [torch.empty_like(Ws[0]) for _ in range(nbuf)]
    out = {}
    for mode in ("cold", "warm"):
        best = 1e9
        for _ in range(rounds):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for k in range(3):
                h.solve_poisson(Ws[k % nbuf], out=Ps[k % nbuf])
            torch.cuda.synchronize()
            e0.record()
            for k in range(reps):
                j = (k % nbuf) if mode == "cold" else 0
                h.solve_poisson(Ws[j], out=Ps[j])
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / reps * 1e3)
        out[mode] = best
    bytes_alg = 32.0 * N * N
    print(f"time N={N:5d} cold {out['cold']:7.2f} us ({bytes_alg / out['cold'] / 1e3:7.1f} GB/s)   "
          f"warm {out['warm']:7.2f} us ({bytes_alg / out['warm'] / 1e3:7.1f} GB/s)", flush=True)
    if N >= 256:
        W = Ws[0] / torch.linalg.norm(Ws[0]) * np.sqrt(N)
        ph = h.profile_iteration(W, 0.25 * qf.hbar(N), reps=10)
        print(f"     in-loop phases (ms): {ph}", flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--sizes", default="256,512,1024,2048")
    ap.add_argument("--check-sizes", default="2,3,5,16,31,32,33,64,127,128,129,200,257,512,1000,1024,1025,1100,2047,2048,2050,3000,4096")
    a = ap.parse_args()
    print("env:", {k: v for k, v in os.environ.items() if k.startswith("QF_")}, flush=True)
    if a.check:
        check([int(x) for x in a.check_sizes.split(",")])
    for N in [int(x) for x in a.sizes.split(",") if x]:
        timeit(N)
