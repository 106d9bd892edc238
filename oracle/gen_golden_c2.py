#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — BASELINE config 2 at full length, from the REAL reference.

    python oracle/gen_golden_c2.py            # build container only (imports /root/reference), about 20 minutes on 8 cores

S(512) = shr2mat(random_shr(lmax=10, s=0, gamma=0, seed=42), 512), dt = 0.25*hbar, tol='auto', maxit=10, minit=1,
10 000 isomp steps in chunks of 1000 (the chunking is part of the fixture: every call re-zeroes the iterate dW,
isospectral.py:430, exactly like qf.solve with steps_out=1000).  Stored in tests/golden/isomp_S_N512_10k.npz:
Casimirs C_2..C_4 and max eigenvalue drift after every chunk, the mean iteration count of every chunk, and of the final
state its Frobenius / infinity norms, a 48 x 48 corner block, a 33-diagonal band and 4096 sampled entries.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from oracle.isomp_oracle import casimirs  # noqa: E402
from oracle.gen_golden import band  # noqa: E402  (also generates nothing on import: guarded by __main__)

qf = refshim.load(with_quantization=True)
import quflow.quantization as quq  # noqa: E402
import quflow.analysis as qua  # noqa: E402
from quflow.geometry import hbar  # noqa: E402
from quflow.integrators.isospectral import isomp_fixedpoint  # noqa: E402


def main():
    N, chunk, nchunks = 512, 1000, 10
    omega = qua.random_shr(lmax=10, s=0.0, gamma=0.0, seed=42)
    W0 = quq.shr2mat(omega, N=N)
    dt = 0.25 * hbar(N)
    ev0 = np.linalg.eigvalsh(1j * W0)
    W = W0.copy()
    cas, its, evd, tols = [casimirs(W0)], [], [0.0], []
    t0 = time.time()
    for c in range(nchunks):
        st = {'iterations': 0.0}
        W = isomp_fixedpoint(W, dt, steps=chunk, stats=st)
        cas.append(casimirs(W))
        its.append(st['iterations'])
        tols.append(st['tol_auto'])
        evd.append(float(np.abs(np.linalg.eigvalsh(1j * W) - ev0).max()))
        print(f"chunk {c + 1}/{nchunks}: {time.time() - t0:.0f}s it/step={st['iterations']:.3f} eig drift={evd[-1]:.3e}", flush=True)
    rng = np.random.RandomState(123)
    idx = rng.randint(0, N, size=(4096, 2))
    out = os.path.join(ROOT, "tests", "golden", "isomp_S_N512_10k.npz")
    np.savez_compressed(out, dt=dt, chunk=chunk, nchunks=nchunks, casimirs=np.array(cas), mean_iterations=np.array(its),
                        tol_auto=np.array(tols), eig_drift=np.array(evd), normF=np.linalg.norm(W), normInf=np.linalg.norm(W, np.inf),
                        Wfinal_block=W[:48, :48].copy(), Wfinal_band=band(W, 16), sample_idx=idx,
                        Wfinal_sample=W[idx[:, 0], idx[:, 1]])
    print("wrote", out, os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    main()
