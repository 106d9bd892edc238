"""Build recipe for libquflow_b200.so (nvcc, sm_100a only, in-tree)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(HERE, "..", "csrc"))
LIB = os.path.join(HERE, "libquflow_b200.so")
SOURCES = ["api.cu", "poisson.cu", "zgemm.cu", "isomp.cu", "comm.cu", "probe.cu", "shr.cu"]


def _nccl_paths():
    """Locate NCCL headers + shared object (torch-bundled wheel first, then system)."""
    cands = []
    try:
        import nvidia.nccl as _n
        base = os.path.dirname(_n.__file__) if getattr(_n, "__file__", None) else list(_n.__path__)[0]
        cands.append((os.path.join(base, "include"), os.path.join(base, "lib")))
    except Exception:
        pass
    cands.append(("/usr/include", "/usr/lib/x86_64-linux-gnu"))
    for inc, lib in cands:
        if os.path.exists(os.path.join(inc, "nccl.h")):
            for name in ("libnccl.so.2", "libnccl.so"):
                if os.path.exists(os.path.join(lib, name)):
                    return inc, lib, name
    return None


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.normpath(os.path.join(HERE, "..", "..", "include", "quflow_b200.h")), __file__]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + srcs
    nccl = _nccl_paths()
    if nccl and any(s.endswith("comm.cu") for s in srcs):
        inc, lib, name = nccl
        cmd += ["-DQF_WITH_NCCL", "-I", inc, "-L", lib, f"-l:{name}", "-Xlinker", f"-rpath={lib}"]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
