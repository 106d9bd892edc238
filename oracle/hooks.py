"""TEST INFRASTRUCTURE ONLY — caller hooks of the isomp loop used by the golden generator and by the tests.

The reference lets the caller run host code inside the step: ``forcing`` (isospectral.py:403-414, 511-520, 594-596),
``strang_splitting`` (:466-467, :602-603), ``callback`` (:550-551) and custom / time-dependent Hamiltonians
(:416-424, :488-491).  The functions below use only ``*``, ``+``, ``-`` and scalars, so the same definitions run on
numpy arrays (reference, oracle, host path) and on torch CUDA tensors (device path).  All of them map
skew-Hermitian matrices to skew-Hermitian matrices, which the whole path assumes.
"""
import math


def forcing_linear(P, W):
    """Autonomous forcing: linear damping plus a stream-function term."""
    return W * (-0.3) + P * 0.05


def forcing_time(P, W, time=0.0):
    """Time-dependent forcing (isomp probes for the ``time`` keyword, isospectral.py:406-412)."""
    return P * (0.1 * math.cos(3.0 * time)) + W * (-0.2)


def strang_damp(h, W):
    """Strang half-step: exact flow of dW/dt = -0.1 W over time h; returns a NEW array (the reference rebinds W)."""
    return W * math.exp(-0.1 * h)


def make_ham_scaled(poisson, factor=0.5):
    """Custom autonomous Hamiltonian: a multiple of the default one (takes W only, so the time probe raises TypeError)."""
    def ham_scaled(W):
        return poisson(W) * factor
    return ham_scaled


def make_ham_time(poisson):
    """Time-dependent Hamiltonian H(W, t) = (1 + 0.1 sin(3 t)) Delta^{-1} W."""
    def ham_time(W, time=0.0):
        return poisson(W) * (1.0 + 0.1 * math.sin(3.0 * time))
    return ham_time


CASES = ("callback", "forcing", "forcing_time", "strang", "ham_scaled", "ham_time", "all")


def case_kwargs(case, poisson):
    """Keyword arguments of one hook case (shared by gen_golden.py, test_oracle.py and the GPU tests)."""
    if case == "callback":
        return dict()
    if case == "forcing":
        return dict(forcing=forcing_linear)
    if case == "forcing_time":
        return dict(forcing=forcing_time, time=0.25)
    if case == "strang":
        return dict(strang_splitting=strang_damp)
    if case == "ham_scaled":
        return dict(hamiltonian=make_ham_scaled(poisson))
    if case == "ham_time":
        return dict(hamiltonian=make_ham_time(poisson), time=0.5)
    if case == "all":
        return dict(hamiltonian=make_ham_time(poisson), time=0.1, forcing=forcing_time, strang_splitting=strang_damp,
                    reinitialize=True, minit=2)
    raise KeyError(case)
