"""TEST INFRASTRUCTURE ONLY — import the real reference in the BUILD container.

``/root/reference`` (klasmodin/quflow) is pure Python but ``import quflow``
fails here because its ``__init__`` star-imports modules that need h5py,
appdirs, ducc0/pyssht and matplotlib, none of which is installed (SURVEY.md
§8c).  This shim registers a bare ``quflow`` package (so ``__init__.py`` never
runs) plus empty stand-ins for the absent third-party modules, after which the
hot-path modules import unmodified.

Used by ``oracle/gen_golden.py`` (fixture generation), by tests that are
skipped when no reference is present, and by ``bench.py``'s reference arm /
``cpu_baseline`` leg, which time the real numba/BLAS ``isomp_fixedpoint``.
On the GPU box ``/root/reference`` does not exist: there the shim loads the
unmodified hot-path modules staged under ``oracle/_ref/`` by
``oracle/make_ref.py`` (git-ignored, shipped with the snapshot).  Nothing on
the product path (``quflow_b200/``) touches this.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")     # oracle/make_ref.py


def _pick_root() -> str:
    env = os.environ.get("QUFLOW_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/quflow"):
        return "/root/reference"
    return _STAGED          # the GPU box: only the staged hot-path modules exist (no quantization / analysis)


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "quflow"))


def full_tree() -> bool:
    """True when the whole reference tree is present (build container), not just the staged hot-path modules."""
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "quflow", "quantization.py"))


def load(with_quantization: bool = False):
    """Return the reference ``quflow`` package object (hot-path modules imported)."""
    if not available():
        raise ImportError(f"reference tree not found at {REFERENCE_ROOT}")
    if "quflow" not in sys.modules or not getattr(sys.modules["quflow"], "_qf_shim", False):
        pkg = types.ModuleType("quflow")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "quflow")]
        pkg._qf_shim = True
        sys.modules["quflow"] = pkg
    import quflow.geometry  # noqa: F401
    import quflow.laplacian  # noqa: F401
    import quflow.integrators  # noqa: F401
    if with_quantization:
        for name, attrs in (("h5py", {"File": object, "Dataset": object, "Group": object}),
                            ("appdirs", {"user_data_dir": lambda *a, **k: "/nonexistent/quflow"}),
                            ("ducc0", {})):
            if name not in sys.modules:
                try:
                    __import__(name)
                except ImportError:
                    mod = types.ModuleType(name)
                    for k, v in attrs.items():
                        setattr(mod, k, v)
                    sys.modules[name] = mod
        os.environ.setdefault("QUFLOW_SAVE_COMPUTED_BASIS", "0")
        import quflow.io  # noqa: F401
        import quflow.quantization  # noqa: F401
        import quflow.transforms  # noqa: F401
        import quflow.analysis  # noqa: F401
    return sys.modules["quflow"]
