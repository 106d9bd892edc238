"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/*.h declares,
the Python mirror keeps the reference's signatures, and the product path fails loudly without a GPU."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from quflow_b200._cuda import build, binding
    build.build()                      # nvcc cross-compiles sm_100a without a GPU
    return binding.library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "quflow_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qf_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from quflow_b200._cuda import binding
    syms = declared_symbols()
    assert len(syms) >= 18
    for name in syms:
        assert hasattr(lib, name), f"{name} declared in include/quflow_b200.h but not exported"
        assert name in binding.SYMBOLS, f"{name} has no ctypes prototype in binding.py"


def test_no_torch_types_in_the_abi():
    text = open(os.path.join(ROOT, "include", "quflow_b200.h")).read()
    code = re.sub(r"/\*.*?\*/", "", text, flags=re.S)      # signatures only, comments stripped
    assert "torch" not in code.lower() and "at::" not in code and "c10::" not in code and "std::" not in code
    assert "#include <stdint.h>" in code and code.count("#include") == 1


def test_library_reports_version_and_fails_loudly_without_gpu(lib):
    import torch
    assert b"quflow_b200" in lib.qf_version()
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import quflow_b200 as qf
    from quflow_b200._cuda import QfError
    W = np.zeros((8, 8), dtype=np.complex128)
    for call in (lambda: qf.solve_poisson(W), lambda: qf.laplace(W), lambda: qf.isomp(W, 0.1, 1)):
        with pytest.raises(QfError, match="no CUDA device|no CPU fallback"):
            call()


def test_product_path_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "quflow_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{f} imports the oracle"
                assert "libqforacle" not in src


def test_python_signatures_mirror_the_reference():
    import quflow_b200 as qf
    spec = inspect.getfullargspec(qf.isomp)
    # quflow/integrators/isospectral.py:338-353
    assert spec.args == ['W', 'dt', 'steps', 'hamiltonian', 'time', 'forcing', 'strang_splitting', 'stats', 'callback',
                         'tol', 'maxit', 'minit', 'verbatim', 'compsum', 'reinitialize']
    defaults = dict(zip(spec.args[-len(spec.defaults):], spec.defaults))
    assert defaults['steps'] == 100 and defaults['tol'] == 'auto' and defaults['maxit'] == 10 and defaults['minit'] == 1
    assert defaults['hamiltonian'] is qf.solve_poisson and defaults['compsum'] is False and defaults['reinitialize'] is False
    assert qf.isomp is qf.isomp_fixedpoint                           # isospectral.py:617
    assert 'stats' in spec.args                                      # qf.solve() passes stats only if it is listed (simulation.py:730)
    assert inspect.getfullargspec(qf.solve_poisson).args == ['W', 'reduce']   # cpu.py:681; no `time` => autonomous probe
    assert qf.hbar(512) == 2.0 / np.sqrt(512.0 ** 2 - 1)             # geometry.py:7-9


def test_sharded_integrator_and_quantization_signatures():
    """The multi-GPU integrator object lists the reference's isomp parameters (so that solve() finds `stats` by
    introspection, quflow/simulation.py:729, like it does for the reference's IsompCUDA object); mat2shr / shr2mat keep the
    reference's names and argument meaning plus the explicit basis (quflow/quantization.py:440-519)."""
    import quflow_b200 as qf
    from quflow_b200.distributed import ShardedIsomp
    spec = inspect.getfullargspec(ShardedIsomp.__call__)
    assert spec.args[1:] == inspect.getfullargspec(qf.isomp).args
    assert ShardedIsomp.device_resident is True
    assert inspect.getfullargspec(qf.mat2shr).args == ['W', 'basis', 'elmax']
    assert inspect.getfullargspec(qf.shr2mat).args == ['omega', 'basis', 'N']
    from quflow_b200.quantization import basis_size
    assert basis_size(33) == sum((33 - m) ** 2 for m in range(33))      # host arithmetic only: no CUDA call


def test_logger_signatures_mirror_the_reference():
    """quflow/geometry.py:53-76 and quflow/physics.py:9-38: same names and argument lists."""
    import quflow_b200 as qf
    sig = lambda f: inspect.getfullargspec(f).args   # noqa: E731
    assert sig(qf.inner_L2) == ['P', 'W'] and sig(qf.norm_L2) == ['W']
    assert sig(qf.energy_euler) == ['W'] and sig(qf.enstrophy) == ['W']
    assert sig(qf.inner_Hm1) == ['W1', 'W2'] and sig(qf.norm_Hm1) == ['W']
    assert sig(qf.inner_H1) == ['P1', 'P2'] and sig(qf.norm_H1) == ['P']
    assert qf.physics.energy_euler is qf.energy_euler


def test_argument_validation_happens_before_any_device_work():
    import quflow_b200 as qf
    W = np.zeros((8, 8), dtype=np.complex128)
    with pytest.raises(AssertionError, match="minit must be at least 1"):
        qf.isomp(W, 0.1, 1, minit=0)
    with pytest.raises(AssertionError, match="maxit must be at minit"):
        qf.isomp(W, 0.1, 1, minit=5, maxit=2)
    with pytest.raises(NotImplementedError):                       # multi-state input with hooks is not part of the path
        qf.isomp(np.zeros((2, 8, 8), dtype=np.complex128), 0.1, 1, callback=lambda W, dW: None)
    with pytest.raises(TypeError):
        qf.solve_poisson(W.astype(np.complex64))
    with pytest.raises(TypeError):
        qf.solve_poisson(W, time=0.0)


@pytest.mark.parametrize("N", [2, 3, 5, 33, 96, 127, 512, 1000, 1024, 2047, 2048, 2050, 3000, 4096, 8192, 10000, 16384])
def test_poisson_work_plan_covers_the_triangle_exactly_once(N):
    """Host logic of the Poisson kernel (csrc/poisson.cu: qf_poisson_plan_host): every element (k, k+m) of the upper
    triangle belongs to exactly one (unit, diagonal slot, local position); pieces do not overlap inside a unit; a band
    longer than one CTA sits on consecutive, cluster-aligned ranks; the units are (almost) full."""
    from quflow_b200._cuda import binding
    plan = binding.poisson_plan(N)
    assert plan is not None
    p, units = plan
    L, M, NT, CL, PC, n = (p[k] for k in ("L", "M", "NT", "CL", "PC", "nunits"))
    assert PC == (NT // M) * L and NT % 32 == 0 and NT <= 512 and n == len(units) and n % CL == 0
    nbands = (N + M - 1) // M
    covered = np.zeros(nbands, dtype=np.int64)           # positions of slot 0 covered per band
    seen_long = {}
    for u, (bL, posbase, bS, PS, nlink, rank0, *_rest) in enumerate(units):
        rank = u % CL
        assert 0 <= PS <= PC
        if bL >= 0:
            lenL = N - M * bL
            piece = min(PC, lenL - posbase)
            assert piece > 0 and posbase % PC == 0 and piece <= PS
            covered[bL] += piece
            if nlink > 1:                                # linked band: group rank g holds positions [g PC, (g+1) PC)
                g = rank - rank0
                assert 0 <= g < nlink and posbase == g * PC and rank0 + nlink <= CL and nlink == -(-lenL // PC)
                seen_long.setdefault(bL, []).append((u, g))
            else:
                assert posbase == 0 and lenL <= PC
        if bS >= 0:
            lenS = N - M * bS
            assert bL >= 0 and PS + lenS <= PC             # the short band fits behind the long piece
            covered[bS] += lenS
    for b in range(nbands):
        assert covered[b] == N - M * b, f"band {b} of N={N}"
    for b, lst in seen_long.items():                     # consecutive units of ONE cluster, group ranks 0..k-1
        us = [u for u, _ in lst]
        assert us == list(range(us[0], us[0] + len(us))) and us[0] // CL == us[-1] // CL
        assert [g for _, g in lst] == list(range(len(lst)))
    if N >= 512 and N % 64 == 0:
        used = sum(N - M * b for b in range(nbands))
        # the folding and the group packing keep the CTAs full; beyond N = 4096 there are too few short bands to fill
        # the last rank of every long band (on average half empty): 79-89 % there
        assert used / (n * PC) > (0.9 if CL <= 2 else 0.75)
    assert binding.poisson_plan(20000) is None           # beyond 8 CTAs per band: fallback kernel
