#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY — generate tests/golden/*.npz from the REAL reference.

Run in the build container (where /root/reference exists):

    python oracle/gen_golden.py

It imports the unmodified reference through ``oracle/refshim.py`` and stores
inputs/outputs of its isomp hot path as small fixtures.  The fixtures travel to
the GPU box (the reference tree does not), where they pin both the CPU oracle
(``tests/test_oracle.py``) and the CUDA path (``tests/test_*_gpu.py``).

Fixture inventory is documented in tests/golden/README.md.
"""
import importlib.util
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402
from oracle.isomp_oracle import random_skewherm, casimirs  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

qf = refshim.load(with_quantization=True)
import quflow.laplacian.cpu as qucpu  # noqa: E402
import quflow.laplacian.gpu as qulegacy  # noqa: E402
import quflow.quantization as quq  # noqa: E402
import quflow.analysis as qua  # noqa: E402
import quflow.utils as quu  # noqa: E402
from quflow.geometry import hbar  # noqa: E402
from quflow.integrators.isospectral import isomp_fixedpoint  # noqa: E402


def save(name, **arrays):
    path = os.path.join(OUT, name)
    np.savez_compressed(path, **arrays)
    print(f"  wrote {name}  ({os.path.getsize(path) / 1024:.0f} KiB)")


class CountingHamiltonian:
    """Wraps the reference solve_poisson to count fixed-point iterations per step.

    Takes only ``W`` so that isomp's ``hamiltonian(W, time=...)`` probe raises
    TypeError and the run is autonomous, exactly like the default path."""

    def __init__(self):
        self.calls = 0

    def __call__(self, W):
        self.calls += 1
        return qucpu.solve_poisson(W)


def run_reference(W0, dt, steps, snapshots=(), **kw):
    """Run the reference isomp once for `steps`; return final W, per-step iteration
    counts, stats and snapshots of W after the requested step numbers."""
    ham = CountingHamiltonian()
    per_step, snaps, last = [], {}, [0]

    def callback(W, dW):
        # called once per step just before W += dW  (isospectral.py:550-551)
        per_step.append(ham.calls - last[0])
        last[0] = ham.calls
        step_no = len(per_step)
        if step_no in snapshots:
            snaps[step_no] = (W + dW).copy()

    stats = {'iterations': 0.0}
    W = isomp_fixedpoint(W0.copy(), dt, steps=steps, hamiltonian=ham, stats=stats, callback=callback, **kw)
    return W, np.array(per_step, dtype=np.int32), stats, snaps


# ---------------------------------------------------------------------------
# A. the reference's own golden vector (tests/test_integrators.py:58-319)
# ---------------------------------------------------------------------------
def gen_reference_golden():
    print("A. reference golden vector N=16")
    spec = importlib.util.spec_from_file_location(
        "ref_test_integrators", os.path.join(refshim.REFERENCE_ROOT, "tests", "test_integrators.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["quflow"].hbar = hbar  # the test module does `import quflow as qf`
    spec.loader.exec_module(mod)
    W0, Wfinal, stepsize, steps = mod.get_isomp_reference_solution()
    dt = hbar(W0.shape[-1]) * stepsize
    W_legacy = isomp_fixedpoint(W0.copy(), dt, steps, hamiltonian=qulegacy.solve_poisson)
    W_head = isomp_fixedpoint(W0.copy(), dt, steps)
    print("   |legacy - Wfinal|max =", np.abs(W_legacy - Wfinal).max(),
          "  |head - Wfinal|max =", np.abs(W_head - Wfinal).max())
    save("ref_isomp_golden_N16.npz", W0=W0, Wfinal=Wfinal, stepsize=stepsize, steps=steps,
         W_head=W_head, W_legacy=W_legacy)


# ---------------------------------------------------------------------------
# B. Poisson known answers (tests/test_laplacian.py:46-71, 134-152, 226-252)
# ---------------------------------------------------------------------------
def poisson_exact(N, seed, zerotrace):
    np.random.seed(seed)
    omegaP = np.random.randn(N ** 2)
    omegaW = omegaP.copy()
    ells = quu.ind2elm(np.arange(N ** 2))[0][1:]
    omegaW[1:] *= -ells * (ells + 1)
    if zerotrace:
        omegaW[0] = 0.0
    omegaP[0] = 0.0
    return quq.shr2mat(omegaP, N=N), quq.shr2mat(omegaW, N=N)


def gen_poisson():
    print("B. Poisson exact solutions + reference outputs")
    for N, zt in ((33, True), (33, False), (64, True), (64, False), (101, True)):
        Pexact, Wexact = poisson_exact(N, seed=N, zerotrace=zt)
        P_ref = qucpu.solve_poisson(Wexact).copy()
        W_ref = qucpu.laplace(Pexact).copy()
        assert np.abs(P_ref - Pexact).max() < 1e-14 * N ** 2
        save(f"poisson_exact_N{N}_zt{int(zt)}.npz", Pexact=Pexact, Wexact=Wexact, P_ref=P_ref, W_ref=W_ref)
    # white-noise input (all diagonals populated)
    for N in (64, 127):
        W = random_skewherm(N, seed=7)
        save(f"poisson_random_N{N}.npz", W=W, P_ref=qucpu.solve_poisson(W).copy(),
             lap=qucpu.laplacian(N, bc=True).copy() if N == 64 else np.zeros(0))


# ---------------------------------------------------------------------------
# C. isomp trajectories
# ---------------------------------------------------------------------------
def gen_isomp_random():
    print("C1. isomp on R(N,42), natural mode (stepsize 0.25, tol auto)")
    for N, steps, snaps in ((32, 100, (1, 10)), (64, 100, (1, 10)), (128, 1000, (1, 100))):
        W0 = random_skewherm(N, 42)
        dt = 0.25 * hbar(N)
        t = time.time()
        W, its, stats, sn = run_reference(W0, dt, steps, snapshots=snaps)
        print(f"   N={N} steps={steps}: {time.time() - t:.1f}s  it/step={stats['iterations']:.3f}")
        arrays = dict(W0=W0, Wfinal=W, dt=dt, steps=steps, iterations=its, tol_auto=stats['tol_auto'],
                      mean_iterations=stats['iterations'], number_of_maxit=stats['number_of_maxit'],
                      casimirs0=casimirs(W0), casimirs=casimirs(W))
        for s, Ws in sn.items():
            arrays[f"W_step{s}"] = Ws
        save(f"isomp_R_N{N}.npz", **arrays)

    print("C2. option variants on R(32,42)")
    N = 32
    W0 = random_skewherm(N, 42)
    dt = 0.25 * hbar(N)
    for tag, kw in (("compsum", dict(compsum=True)), ("tol1e-10", dict(tol=1e-10)),
                    ("reinit", dict(reinitialize=True)), ("minit3", dict(minit=3, maxit=5))):
        W, its, stats, _ = run_reference(W0, dt, 50, **kw)
        save(f"isomp_R_N32_{tag}.npz", Wfinal=W, dt=dt, steps=50, iterations=its,
             mean_iterations=stats['iterations'], number_of_maxit=stats['number_of_maxit'],
             tol_auto=stats.get('tol_auto', np.nan))

    print("C3. profile mode (dt = 0.01 hbar, minit = maxit = 10; profiling/run_profiling.py:124-127)")
    N = 64
    W0 = random_skewherm(N, 42)
    dt = 0.01 * hbar(N)
    W, its, stats, _ = run_reference(W0, dt, 5, minit=10, maxit=10)
    save("isomp_R_N64_profile.npz", Wfinal=W, dt=dt, steps=5, iterations=its,
         mean_iterations=stats['iterations'], number_of_maxit=stats['number_of_maxit'])


def band(W, lmax):
    N = W.shape[-1]
    out = np.zeros((2 * lmax + 1, N), dtype=W.dtype)
    for m in range(-lmax, lmax + 1):
        d = np.diagonal(W, m)
        out[m + lmax, :d.shape[0]] = d
    return out


def gen_isomp_smooth():
    print("C4. isomp on S(N) = shr2mat(random_shr(lmax=10, s=0, gamma=0, seed=42), N)")
    omega = qua.random_shr(lmax=10, s=0.0, gamma=0.0, seed=42)
    for N in (64, 512):
        t = time.time()
        W0 = quq.shr2mat(omega, N=N)
        dt = 0.25 * hbar(N)
        W, its, stats, sn = run_reference(W0, dt, 100, snapshots=(1,))
        print(f"   N={N}: {time.time() - t:.1f}s  it/step={stats['iterations']:.3f}  "
              f"L2={np.linalg.norm(W0) / np.sqrt(N)!r} Linf={np.linalg.norm(W0, np.inf)!r}")
        arrays = dict(omega=omega, W0_band=band(W0, 10), dt=dt, steps=100, iterations=its,
                      tol_auto=stats['tol_auto'], mean_iterations=stats['iterations'],
                      normF0=np.linalg.norm(W0), normInf0=np.linalg.norm(W0, np.inf), specnorm0=np.linalg.norm(W0, 2),
                      normF=np.linalg.norm(W), normInf=np.linalg.norm(W, np.inf),
                      casimirs0=casimirs(W0), casimirs=casimirs(W))
        assert np.abs(W0 - sum(np.diag(np.diagonal(W0, m), m) for m in range(-10, 11))).max() == 0.0
        if N <= 64:
            arrays.update(Wfinal=W, W_step1=sn[1])
        else:
            # N=512: keep the fixture small — a corner block, a band and a fixed random sample
            rng = np.random.RandomState(123)
            idx = rng.randint(0, N, size=(4096, 2))
            arrays.update(Wfinal_block=W[:48, :48].copy(), Wfinal_band=band(W, 16),
                          sample_idx=idx, Wfinal_sample=W[idx[:, 0], idx[:, 1]],
                          W_step1_band=band(sn[1], 16))
        save(f"isomp_S_N{N}.npz", **arrays)


# ---------------------------------------------------------------------------
# H. the callers' hooks of the loop: callback, forcing, strang_splitting, custom / time-dependent Hamiltonians
# ---------------------------------------------------------------------------
def gen_hooks():
    from oracle import hooks
    print("H. isomp with caller hooks on R(32,42), 30 steps")
    N, steps = 32, 30
    W0 = random_skewherm(N, 42)
    dt = 0.25 * hbar(N)
    arrays = dict(W0=W0, dt=dt, steps=steps)
    for case in hooks.CASES:
        kw = hooks.case_kwargs(case, lambda W: qucpu.solve_poisson(W).copy())
        per_step, cb_norms = [], []
        calls, last = [0], [0]
        ham = kw.pop("hamiltonian", None)
        if ham is None:
            def counted(W):
                calls[0] += 1
                return qucpu.solve_poisson(W)
        elif "time" in ham.__code__.co_varnames:
            def counted(W, time=0.0, _h=ham):
                calls[0] += 1
                return _h(W, time=time)
        else:
            def counted(W, _h=ham):
                calls[0] += 1
                return _h(W)

        def callback(W, dW):
            per_step.append(calls[0] - last[0])
            last[0] = calls[0]
            cb_norms.append((np.linalg.norm(W), np.linalg.norm(dW)))

        stats = {'iterations': 0.0}
        W = isomp_fixedpoint(W0.copy(), dt, steps=steps, hamiltonian=counted, stats=stats, callback=callback, **kw)
        its = np.array(per_step, dtype=np.int32)
        if "time" in kw and ham is not None:
            its[0] -= 1          # the autonomy probe hamiltonian(W, time=time) of isospectral.py:419-421
        print(f"   {case:13s} it/step={stats['iterations']:.3f}  |W|={np.linalg.norm(W):.6f}")
        arrays.update({f"{case}_Wfinal": W, f"{case}_iterations": its, f"{case}_tol_auto": stats['tol_auto'],
                       f"{case}_mean_iterations": stats['iterations'], f"{case}_cb_norms": np.array(cb_norms)})
    save("isomp_hooks_N32.npz", **arrays)


# ---------------------------------------------------------------------------
# M. multi-state (k, N, N): members 1.. advected by member 0's stream function (cpu.py:672-674)
# ---------------------------------------------------------------------------
def gen_multistate():
    print("M. multi-state isomp, k=3 members of R(32, seed), 30 steps")
    N, steps = 32, 30
    W0 = np.stack([random_skewherm(N, s) for s in (42, 43, 44)])
    dt = 0.25 * hbar(N)
    ham = CountingHamiltonian()
    per_step, last = [], [0]

    def callback(W, dW):
        per_step.append(ham.calls - last[0])
        last[0] = ham.calls

    arrays = dict(W0=W0, dt=dt, steps=steps)
    for tag, kw in (("plain", dict()), ("compsum", dict(compsum=True))):
        del per_step[:]
        ham.calls = last[0] = 0
        stats = {'iterations': 0.0}
        W = isomp_fixedpoint(W0.copy(), dt, steps=steps, hamiltonian=ham, stats=stats, callback=callback, **kw)
        print(f"   {tag}: it/step={stats['iterations']:.3f}")
        arrays.update({f"{tag}_Wfinal": W, f"{tag}_iterations": np.array(per_step, dtype=np.int32),
                       f"{tag}_tol_auto": stats['tol_auto'], f"{tag}_mean_iterations": stats['iterations']})
    save("isomp_multistate_N32.npz", **arrays)


if __name__ == "__main__":
    parts = dict(A=gen_reference_golden, B=gen_poisson, C=gen_isomp_random, S=gen_isomp_smooth, H=gen_hooks,
                 M=gen_multistate)
    for key in (sys.argv[1:] or list(parts)):
        parts[key]()
    print("done")
