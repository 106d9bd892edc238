"""TEST INFRASTRUCTURE ONLY — CPU oracle for the isomp hot path.

Nothing under ``quflow_b200/`` may import this package.  Allowed importers:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py`` (cpu_baseline /
``--impl reference`` legs).  Parity status: PINNED against the reference
(see ``oracle/gen_golden.py`` and ``tests/golden/README.md``).
"""
from .isomp_oracle import (  # noqa: F401
    hbar,
    laplacian,
    solve_poisson,
    solve_poisson_numpy,
    laplace,
    conj_subtract_,
    norm_inf,
    isomp_fixedpoint,
    isomp,
    casimirs,
    random_skewherm,
    build_c_oracle,
)
