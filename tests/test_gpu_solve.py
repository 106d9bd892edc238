"""GPU tests of solve() with the device-resident isomp (BASELINE config 3's driver path, without HDF5)."""
import os

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


def test_qusimulation_stores_shr_through_device_mat2shr(cuda_device, tmp_path, monkeypatch):
    """QuSimulation with qutypes {'mat', 'shr'}: the spherical-harmonic coefficients of every record come from the device
    mat2shr (reference layout: dataset 'shr' (T, N**2), attribute qutype, quflow/simulation.py:297-304, 367-375); the run
    goes through solve() with the asynchronous output pipeline."""
    import sys
    import fake_h5py
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    import quflow_b200 as qf
    from oracle import shr_oracle
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "shr_N33.npz"))
    N = 33
    W = np.ascontiguousarray(g["W"])
    fn = str(tmp_path / "shr.hdf5")
    sim = qf.QuSimulation(fn, overwrite=True, state=W, qutypes={'mat': None, 'shr': None}, basis=g["basis"])
    with pytest.raises(ValueError):
        qf.QuSimulation(str(tmp_path / "nobasis.hdf5"), overwrite=True, state=W, qutypes={'shr': None})
    qf.solve(W, stepsize=0.2, steps=30, steps_out=10, callback=sim, progress_bar=False)
    assert sim['mat'].shape == (4, N, N) and sim['shr'].shape == (4, N * N)
    f = fake_h5py.File(fn, "r")
    assert f["/shr"].attrs["qutype"] == "shr" and f["/shr"].dtype == np.float64
    for r in range(4):
        ref = shr_oracle.mat2shr(sim['mat', r], g["basis"])
        assert np.abs(sim['shr', r] - ref).max() < 1e-13 * np.abs(ref).max()
    np.testing.assert_array_equal(sim['mat', -1], W)          # the caller's array ends up advanced, as with the reference
    again = qf.QuSimulation(fn, basis=g["basis"])             # reopening a file with 'shr' needs the basis again
    assert set(again.qutypes) == {'mat', 'shr'}
    with pytest.raises(ValueError):
        qf.QuSimulation(fn)
