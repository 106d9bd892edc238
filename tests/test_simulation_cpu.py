"""CPU tests of the callers around the hot path: solve() bookkeeping and the QuSimulation HDF5 layout
(reference: quflow/simulation.py; upstream tests/test_simulation.py:130-168).  h5py is absent from this image, so the
store is exercised against a small in-memory stand-in (tests/fake_h5py.py); the integrator is a cheap CPU stand-in
because these tests are about the driver, not the kernels (the GPU path of solve() is in tests/test_gpu_solve.py)."""
import pickle
import sys

import numpy as np
import pytest

import fake_h5py


@pytest.fixture()
def sim_mod(monkeypatch):
    monkeypatch.setitem(sys.modules, "h5py", fake_h5py)
    import quflow_b200.simulation as sm
    return sm


def rand_skew(N, seed=0):
    rng = np.random.RandomState(seed)
    A = rng.randn(N, N) + 1j * rng.randn(N, N)
    return np.ascontiguousarray(A - A.conj().T)


def fake_integrator(W, dt, steps=100, hamiltonian=None, time=None, stats=None, callback=None, tol='auto'):
    """In-place 'integrator' with the isomp calling convention: rotates W by a phase per step."""
    W *= np.exp(1j * dt * steps)
    if stats:
        stats['iterations'] = 3.0
        stats['tol_auto'] = 1e-8
        stats['number_of_maxit'] = 0.0
    return W


def test_solve_time_and_step_bookkeeping(sim_mod, tmp_path):
    import quflow_b200 as qf
    N = 12
    W = rand_skew(N)
    sim = sim_mod.QuSimulation(str(tmp_path / "a.hdf5"), overwrite=True, state=W,
                               loggers={'norm': lambda X: np.linalg.norm(X)})
    sim_mod.solve(W, stepsize=0.1, steps=100, steps_out=10, progress_bar=False, callback=sim, integrator=fake_integrator)
    np.testing.assert_allclose(0.1 * qf.hbar(N) * 10 * np.arange(11), sim['time'])      # upstream :141
    np.testing.assert_equal(10 * np.arange(11), sim['step'])                            # upstream :142
    assert sim['norm', -1] == np.linalg.norm(sim['mat', -1])
    assert sim['mat'].shape == (11, N, N)
    np.testing.assert_equal(sim['iterations'], [0.0] + [3.0] * 10)                      # stats forwarded to the store
    np.testing.assert_equal(sim['mat', -1], W)                                          # caller's array was advanced in place


def test_solve_argument_rules(sim_mod):
    W = rand_skew(8)
    seen = []
    cb = lambda X, delta_time, delta_steps, **kw: seen.append((delta_steps, delta_time))   # noqa: E731
    with pytest.raises(ValueError, match="Either `dt` or `stepsize`"):
        sim_mod.solve(W, steps=10, integrator=fake_integrator, progress_bar=False)
    sim_mod.solve(W, dt=0.5, steps=25, steps_out=10, callback=cb, integrator=fake_integrator, progress_bar=False)
    assert [s for s, _ in seen] == [10, 10, 5]                                          # last chunk is shorter (:784-787)
    seen.clear()
    sim_mod.solve(W, dt=0.5, simtime=10.0, dt_out=2.0, callback=cb, integrator=fake_integrator, progress_bar=False)
    assert [s for s, _ in seen] == [4] * 5                                              # steps = round(simtime/dt), steps_out = round(dt_out/dt)
    with pytest.raises(ValueError, match="smaller than current"):
        sim_mod.solve(W, dt=0.5, endtime=1.0, time=2.0, integrator=fake_integrator, progress_bar=False)
    with pytest.warns(UserWarning):
        sim_mod.solve(W, dt=0.5, steps=2, simtime=1.0, integrator=fake_integrator, progress_bar=False)


def test_qusimulation_layout_matches_reference(sim_mod, tmp_path):
    N = 9
    W = rand_skew(N)
    fn = str(tmp_path / "b.hdf5")
    sim = sim_mod.QuSimulation(fn, overwrite=True, state=W, time=1.5, energy=0.25)
    f = fake_h5py.File(fn, "r")
    g = f["/"]
    assert set(g.attrs) >= {"version", "created", "qutypes", "loggers", "N"}           # simulation.py:134-143, 374
    assert pickle.loads(bytes(g.attrs["qutypes"][0])) == {'mat': None}
    # byte level: the attribute values are exactly what the reference writes — a length-1 numpy bytes array holding
    # pickle.dumps(...) of the qutypes / loggers dict (quflow/simulation.py:136-142), an ISO timestamp, the version string
    assert g.attrs["qutypes"].shape == (1,) and g.attrs["qutypes"].dtype.kind == "S"
    assert bytes(g.attrs["qutypes"][0]) == np.array([pickle.dumps({'mat': None})])[0]
    assert bytes(g.attrs["loggers"][0]) == np.array([pickle.dumps({})])[0]
    import datetime
    datetime.datetime.fromisoformat(g.attrs["created"])
    assert isinstance(g.attrs["version"], str)
    assert g.attrs["N"] == N
    mat = f["/mat"]
    assert mat.shape == (1, N, N) and mat.chunks == (1, N, N) and mat.maxshape == (None, N, N)   # :364-368
    assert mat.attrs["qutype"] == "mat" and mat.dtype == np.complex128
    assert f["/time"].dtype == np.float64 and f["/time"][0] == 1.5
    assert f["/step"][0] == 0
    for name in ("tol_auto", "iterations", "number_of_maxit", "energy"):                # :409-412, user fields
        assert f["/" + name].shape == (1,)
    assert "args" in g
    sim(W * 2, delta_time=0.5, delta_steps=7, iterations=3.0, unknown_field=1.0)
    assert sim['time', -1] == 2.0 and sim['step', -1] == 7 and sim['iterations', -1] == 3.0
    np.testing.assert_equal(sim[-1], W * 2)                                             # bare index means 'mat' (:246-249)
    sim['stepsize'] = 0.1
    sim['integrator'] = fake_integrator
    assert dict(sim.args())['stepsize'] == 0.1
    assert sim['integrator'] is fake_integrator                                         # pickled callables round-trip
    again = sim_mod.QuSimulation(fn)
    assert again.qutypes == {'mat': None} and again['mat'].shape == (2, N, N)
    with pytest.raises(ValueError):
        sim_mod.QuSimulation(fn, state=W)
    with pytest.raises(NotImplementedError):
        sim_mod.QuSimulation(str(tmp_path / "c.hdf5"), state=W, qutypes={'mat': None, 'fun': np.float32})
    with pytest.raises(ValueError):
        sim_mod.QuSimulation(str(tmp_path / "d.hdf5"), state=W, datapath="/x")


def test_solve_continues_from_a_stored_simulation(sim_mod, tmp_path):
    N = 10
    W = rand_skew(N, 3)
    fn = str(tmp_path / "e.hdf5")
    sim = sim_mod.QuSimulation(fn, overwrite=True, state=W)
    sim['stepsize'] = 0.2
    sim['steps'] = 20
    sim['steps_out'] = 5
    sim['integrator'] = fake_integrator
    sim_mod.solve(sim, progress_bar=False)
    assert list(sim['step']) == [0, 5, 10, 15, 20]
    sim_mod.solve(sim_mod.QuSimulation(fn), progress_bar=False)                         # restart: continues in time
    assert list(sim['step']) == [0, 5, 10, 15, 20, 25, 30, 35, 40]
    import quflow_b200 as qf
    np.testing.assert_allclose(sim['time'], 0.2 * qf.hbar(N) * np.array(sim['step']))
