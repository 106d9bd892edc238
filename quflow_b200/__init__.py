"""quflow_b200 — the isomp hot path of klasmodin/quflow on NVIDIA B200 (sm_100a).

Public names follow the reference package (``import quflow as qf``):
``qf.isomp`` / ``qf.isomp_fixedpoint`` (quflow/integrators/isospectral.py), ``qf.solve_poisson`` /
``qf.laplace`` (quflow/laplacian/cpu.py), ``qf.hbar`` / ``qf.inner_L2`` / ``qf.norm_L2`` (quflow/geometry.py),
``qf.energy_euler`` / ``qf.enstrophy`` / ``qf.inner_Hm1`` / ``qf.inner_H1`` (quflow/physics.py), ``qf.solve`` /
``qf.QuSimulation`` (quflow/simulation.py).  Everything computes in the CUDA library
``quflow_b200/_cuda/libquflow_b200.so``; there is no CPU fallback.
"""
from .geometry import hbar, inner_L2, norm_L2  # noqa: F401
from .physics import energy_euler, enstrophy, inner_Hm1, norm_Hm1, inner_H1, norm_H1  # noqa: F401
from .laplacian import solve_poisson, laplace, select_first  # noqa: F401
from .integrators import isomp, isomp_fixedpoint, isomp_ensemble  # noqa: F401
from .simulation import solve, QuSimulation  # noqa: F401
from .quantization import mat2shr, shr2mat  # noqa: F401
from . import integrators, simulation, physics, quantization, _cuda  # noqa: F401

__version__ = "0.1.0"
