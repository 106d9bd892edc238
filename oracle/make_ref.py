"""TEST INFRASTRUCTURE ONLY — recipe that stages the UNMODIFIED reference hot path under ``oracle/_ref/``.

The reference (klasmodin/quflow) is pure Python: "building" it is copying the handful of modules the isomp path
imports — ``quflow/geometry.py``, ``quflow/integrators/*.py``, ``quflow/laplacian/*.py`` — byte for byte from
``/root/reference`` into ``oracle/_ref/quflow/``.  ``oracle/_ref/`` is git-ignored (no reference source enters the
history) but NOT gpurun-ignored, so the staged copy travels to the GPU box, where ``/root/reference`` does not exist.
The package ``__init__`` of the reference is NOT staged: it star-imports modules that need h5py / ducc0 / matplotlib
(absent from this image); ``oracle/refshim.py`` registers a bare ``quflow`` package instead, so the
staged modules import unmodified (they only need numpy, scipy and numba, all in the image).

Consumers: ``bench.py --impl reference`` and the ``cpu_baseline`` leg (they time the real numba/BLAS
``isomp_fixedpoint``, kind "reference"), and tests that cross-check the oracle port against it.  Nothing under
``quflow_b200/`` may import it.

    python oracle/make_ref.py          # (re)stage; called by __graft_entry__.build() when /root/reference exists
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("QUFLOW_REFERENCE_SRC", "/root/reference")
DEST = os.path.join(HERE, "_ref")
# the modules the isomp path imports (integrators/__init__ pulls erk.py and mhd.py, laplacian/__init__ every backend)
FILES = ["quflow/geometry.py"] + \
        ["quflow/integrators/" + f for f in ("__init__.py", "isospectral.py", "erk.py", "mhd.py")] + \
        ["quflow/laplacian/" + f for f in ("__init__.py", "cpu.py", "direct.py", "gpu.py", "sparse.py", "tridiagonal.py")]


def source_available() -> bool:
    return all(os.path.isfile(os.path.join(REF_SRC, f)) for f in FILES)


def staged() -> bool:
    return os.path.isfile(os.path.join(DEST, "MANIFEST.json")) and all(os.path.isfile(os.path.join(DEST, f)) for f in FILES)


def build_ref(force: bool = False) -> str:
    """Stage the reference modules; returns the staging root.  No-op when already staged from identical bytes."""
    if not source_available():
        if staged():
            return DEST
        raise FileNotFoundError(f"reference tree not found at {REF_SRC} and nothing staged under {DEST}")
    manifest = {}
    for f in FILES:
        with open(os.path.join(REF_SRC, f), "rb") as fh:
            manifest[f] = hashlib.sha256(fh.read()).hexdigest()
    mpath = os.path.join(DEST, "MANIFEST.json")
    if not force and staged():
        try:
            if json.load(open(mpath))["sha256"] == manifest:
                return DEST
        except Exception:
            pass
    for f in FILES:
        dst = os.path.join(DEST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, f), dst)
    with open(mpath, "w") as fh:
        json.dump({"source": REF_SRC, "note": "unmodified copies; quflow/__init__.py deliberately not staged", "sha256": manifest},
                  fh, indent=1)
    return DEST


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
