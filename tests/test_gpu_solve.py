"""GPU tests of solve() with the device-resident isomp (BASELINE config 3's driver path, without HDF5)."""
import numpy as np
import pytest

import oracle
from conftest import relfro

pytestmark = pytest.mark.gpu


def test_solve_device_resident_matches_chunked_reference(cuda_device):
    import quflow_b200 as qf
    N = 64
    W0 = oracle.random_skewherm(N, 5)
    records = []
    W = W0.copy()
    ret = qf.solve(W, stepsize=0.25, steps=35, steps_out=10, progress_bar=False,
                   callback=lambda X, delta_time, delta_steps, **st: records.append((X.copy(), delta_steps, dict(st))))
    assert ret is W
    assert [r[1] for r in records] == [10, 10, 10, 5]
    assert set(records[0][2]) == {'iterations', 'tol_auto', 'number_of_maxit'}
    # the reference semantics: each output interval is a separate integrator call (dW warm start reset per call)
    Wref = W0.copy()
    dt = 0.25 * oracle.hbar(N)
    for n, rec in zip((10, 10, 10, 5), records):
        Wref = oracle.isomp(Wref, dt, n)
        assert relfro(rec[0], Wref) < 1e-12
    assert relfro(W, Wref) < 1e-12


def test_solve_restart_equivalence(cuda_device):
    """upstream tests/test_simulation.py:147-168: 50 + 50 steps equals 100 steps when chunked identically."""
    import quflow_b200 as qf
    N = 35
    W0 = oracle.random_skewherm(N, 8)
    Wa = W0.copy()
    qf.solve(Wa, stepsize=0.1, steps=50, steps_out=10, progress_bar=False)
    qf.solve(Wa, stepsize=0.1, steps=50, steps_out=10, progress_bar=False)
    Wb = W0.copy()
    qf.solve(Wb, stepsize=0.1, steps=100, steps_out=10, progress_bar=False)
    assert np.array_equal(Wa, Wb)


def test_solve_passes_hooks_with_the_numpy_convention(cuda_device):
    """solve() forwards forcing / integrator_callback / time to the integrator (simulation.py:726-733, 788): the hooks
    receive numpy arrays when W is numpy, and `time` advances across output chunks for a time-dependent forcing."""
    import quflow_b200 as qf
    from oracle import hooks
    N = 40
    W0 = oracle.random_skewherm(N, 11)
    dt = 0.2 * oracle.hbar(N)
    seen = []

    def icb(W, dW):
        assert isinstance(W, np.ndarray) and isinstance(dW, np.ndarray)
        seen.append(float(np.linalg.norm(dW)))

    W = W0.copy()
    qf.solve(W, dt=dt, steps=12, steps_out=5, progress_bar=False, forcing=hooks.forcing_time, integrator_callback=icb, time=0.3)
    Wref, t, ref_seen = W0.copy(), 0.3, []
    for n in (5, 5, 2):
        Wref = oracle.isomp(Wref, dt, n, forcing=hooks.forcing_time, time=t,
                            callback=lambda W, dW: ref_seen.append(float(np.linalg.norm(dW))))
        t += n * dt
    assert relfro(W, Wref) < 1e-12
    np.testing.assert_allclose(seen, ref_seen, rtol=1e-11)
