#!/usr/bin/env python
"""bench.py — isomp steps/s on B200 (BASELINE.json metric), roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 2048] [--mode natural|profile]

Workload (config.workload): R(N, 42) — random skew-Hermitian trace-free complex128 vorticity normalised to
norm_L2 = 1, N = 2048 by default (the size BASELINE.json's metric is quoted on; it fits one GPU),
natural mode: dt = 0.25*hbar(N), tol='auto', maxit=10, minit=1 (about 3 fixed-point iterations per step).
A "step" is one isospectral-midpoint time step.

* ``value``  : steps/s with W resident in HBM, timed with CUDA events over exactly K steps (max over ranks);
               ``repeat_values`` times the same region again (run-to-run spread, not part of ``value``).
* ``e2e``    : the same metric through the public Python API with HOST buffers (numpy in pinned memory): every step
               is one ``qf.isomp(W_host, dt, steps=1)`` call = H2D copy of W + one step + D2H copy of W.  On several
               GPUs the host state is row-distributed (``host_rows="own"``): every rank copies its own 1/G of the rows
               in and out over its own PCIe link and the state is completed over NVLink.  ``e2e.copies`` times the
               two copies alone and reports the iterations per step of the one-step calls (each call starts from
               dW = 0 like the reference, so it needs about one more iteration than a step inside a long call).
* ``roofline``: the dominant kernel (k_zgemm3m_ws, FP64 DMMA) — EXECUTED flops per launch / CUDA-event launch time,
               against the FP64 tensor peak MEASURED IN THE SAME RUN (qf_measure_fp64_tensor_peak; MEASURED_PEAKS.json
               has no FP64 entry).  ``roofline_poisson`` reports the HBM-bound Poisson solve against
               MEASURED_PEAKS.json's copy bandwidth.  ``incumbent``: cuBLAS ZGEMM and cusparseDgtsv2StridedBatch, the
               library kernels the reference's own GPU prototype calls, timed in the same run.
* ``parity`` : N=1: the GPU against the CPU arm on the same two calls; N>1: the sharded handle against a single-GPU
               handle (rel. Frobenius, iteration counts, ranks bit-identical by checksum all-gather).
* ``cpu_baseline`` / ``--impl reference``: the UNMODIFIED reference ``isomp_fixedpoint`` (numba Thomas solve + BLAS
               zgemm) staged under oracle/_ref, all host cores (thread counts forced before numpy loads and reported
               through threadpoolctl), on a bounded sample of the same workload; falls back to the oracle port.
* ``--workload ensemble``: BASELINE config 5 (independent N=256 members sharded per member, no collective).
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def cpu_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return n


def _reference_arm_requested(argv):
    return any(a == "--impl=reference" or (a == "--impl" and i + 1 < len(argv) and argv[i + 1] == "reference")
               for i, a in enumerate(argv))


if _reference_arm_requested(sys.argv):
    # The CPU arm uses every host core.  torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, and the BLAS /
    # OpenMP / numba pools read their sizes when the libraries load, so the counts are forced HERE, before numpy is imported.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS", "NUMBA_NUM_THREADS"):
        os.environ[_v] = str(cpu_threads())

import numpy as np  # noqa: E402

FP64_TENSOR_PEAK_TFLOPS_R01 = 37.15   # round-1 measurement (profiles/r01_fp64_pipes.txt); bench.py now measures it live


def incumbents(N, dev):
    """The library kernels the reference's own GPU prototype calls for this path (quflow/experimental/
    isospectral_cuda.py: cuBLASLt ZGEMM; quflow/experimental/cuda.py:123-166: cusparseDgtsv2StridedBatch), timed on the
    same GPU in the same run.  Comparators only: nothing here is on the product path."""
    import re
    import torch
    out = {}
    try:
        A = torch.randn(N, N, dtype=torch.complex128, device=dev)
        B = torch.randn(N, N, dtype=torch.complex128, device=dev)
        C = torch.empty_like(A)
        for _ in range(2):
            torch.matmul(A, B, out=C)
        best = 1e30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            e0.record()
            torch.matmul(A, B, out=C)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out["cublas_zgemm_ms"] = best
        out["cublas_zgemm_tflops_8n3"] = 8.0 * N ** 3 / (best * 1e-3) / 1e12
        del A, B, C
    except Exception as e:
        out["cublas_zgemm_error"] = str(e)[:200]
    exe = os.path.join(ROOT, "tools", "microbench", "cusparse_gtsv_ref")
    if os.path.exists(exe):
        try:
            txt = subprocess.run([exe, str(N)], capture_output=True, text=True, timeout=120).stdout
            m = re.search(r"N=%d .*?: ([0-9.]+) us" % N, txt)
            if m:
                out["cusparse_gtsv2_strided_batch_x2_us"] = float(m.group(1))
                out["cusparse_note"] = "two library solves (re, im) with the prototype's packing; its pack/unpack kernels not included"
        except Exception as e:
            out["cusparse_error"] = str(e)[:200]
    else:
        out["cusparse_error"] = "tools/microbench/cusparse_gtsv_ref not built"
    return out


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def hbar(N):
    return 2.0 / np.sqrt(float(N) ** 2 - 1.0)


def workload(N, seed=42):
    rng = np.random.RandomState(seed)
    A = rng.randn(N, N) + 1j * rng.randn(N, N)
    W = A - A.conj().T
    W -= np.eye(N) * np.trace(W) / N
    W /= np.linalg.norm(W) / np.sqrt(N)
    return np.ascontiguousarray(W)


def mode_kwargs(mode, N):
    if mode == "profile":   # the reference's own protocol, profiling/run_profiling.py:124-127
        return dict(dt=0.01 * hbar(N), maxit=10, minit=10)
    return dict(dt=0.25 * hbar(N), maxit=10, minit=1)


def ncu_traffic(kernel_prefix, index=0):
    """DRAM bytes (read + write) per launch of a kernel from the committed `ncu --set full` capture
    (profiles/r02_ncu_full_kernels.json, N=2048); None for other sizes or when the file is missing."""
    try:
        rows = [r for r in json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_full_kernels.json")))
                if r["kernel"].startswith(kernel_prefix)]
        r = rows[index]

        def to_bytes(txt):
            val, unit = txt.split()
            return float(val) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
        return to_bytes(r["dram_read"]) + to_bytes(r["dram_write"])
    except Exception:
        return None


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # median of the upper half = clocks under load (idle samples at the edges drag the plain median down)
        s = sorted(sm)
        return {"sm_mhz": float(np.median(s[len(s) // 2:])), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, all host cores
# ----------------------------------------------------------------------------------------------------
def thread_report():
    """Thread counts the CPU arm actually runs with (BLAS / OpenMP pools via threadpoolctl, numba's prange pool)."""
    rep = {}
    try:
        import threadpoolctl
        for lib in threadpoolctl.threadpool_info():
            rep[f"{lib.get('user_api')}:{lib.get('internal_api')}"] = lib.get("num_threads")
    except Exception as e:       # pragma: no cover
        rep["threadpoolctl"] = f"unavailable ({e})"
    try:
        import numba
        rep["numba"] = numba.get_num_threads()
    except Exception:
        pass
    return rep


def force_cpu_threads(n):
    """Run-time counterpart of the environment set-up at the top of this file (cpu_baseline leg of the GPU arm, where
    numpy is already loaded): resize the BLAS/OpenMP pools through threadpoolctl and numba's pool."""
    try:
        import threadpoolctl
        threadpoolctl.threadpool_limits(limits=n)
    except Exception:
        pass
    try:
        import numba
        numba.set_num_threads(min(n, numba.config.NUMBA_NUM_THREADS))
    except Exception:
        pass


def run_cpu_reference(N, mode, steps, warmup, keep=None):
    """Time the UNMODIFIED reference isomp_fixedpoint (numba Thomas solve + BLAS zgemm, quflow/integrators/
    isospectral.py:338-613) staged under oracle/_ref by oracle/make_ref.py.  JIT compilation is excluded by one
    throw-away step at N=64 (numba compiles per signature, not per size).  Returns None when it is not staged."""
    from oracle import refshim, make_ref
    if not (make_ref.staged() or refshim.available()):
        return None
    refshim.load()
    from quflow.integrators.isospectral import isomp_fixedpoint
    kw = mode_kwargs(mode, N)
    isomp_fixedpoint(workload(64), 0.25 * hbar(64), steps=1)          # JIT warm-up
    W = workload(N)
    if warmup > 0:
        W = isomp_fixedpoint(W, kw["dt"], steps=warmup, maxit=kw["maxit"], minit=kw["minit"])
    stats = {'iterations': 0.0}
    t0 = time.perf_counter()
    W = isomp_fixedpoint(W, kw["dt"], steps=steps, maxit=kw["maxit"], minit=kw["minit"], stats=stats)
    dt = time.perf_counter() - t0
    if keep is not None:
        keep["W"] = W
    return steps / dt, dt, stats['iterations']


def run_cpu(N, mode, steps, warmup, keep=None):
    """(value, seconds, iterations/step, kind, description): the real reference when staged, else the oracle port.
    `keep` (a dict) receives the final state, for the parity check of the GPU arm."""
    r = run_cpu_reference(N, mode, steps, warmup, keep)
    if r is not None:
        return r + ("reference", "unmodified reference isomp_fixedpoint (numba prange Thomas solve + numpy BLAS zgemm) from oracle/_ref")
    return run_cpu_port(N, mode, steps, warmup, keep) + ("port", "oracle port: numpy BLAS zgemm + OpenMP Thomas (oracle/)")


def run_cpu_port(N, mode, steps, warmup, keep=None):
    """Time the CPU oracle (port of isospectral.py:338-613 + cpu.py:281-362) on `steps` steps."""
    import oracle
    kw = mode_kwargs(mode, N)
    W = workload(N)
    stats = {'iterations': 0.0}
    if warmup > 0:
        oracle.isomp(W, kw["dt"], steps=warmup, maxit=kw["maxit"], minit=kw["minit"])
    t0 = time.perf_counter()
    oracle.isomp(W, kw["dt"], steps=steps, maxit=kw["maxit"], minit=kw["minit"], stats=stats)
    dt = time.perf_counter() - t0
    if keep is not None:
        keep["W"] = W
    return steps / dt, dt, stats['iterations']


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.n
    cores = cpu_threads()
    force_cpu_threads(cores)     # the environment was already forced before numpy loaded (top of this file)
    # bounded: at N=2048 one CPU step costs about half a second; keep the whole run within a few minutes
    warm = min(args.warmup, 1)
    val, secs, its, kind, what = run_cpu(N, args.mode, args.steps, warm)
    sample = (f"{args.steps} steps (+{warm} warm-up) of R({N},42), {args.mode} mode, {its:.2f} it/step, {secs:.1f} s; {what}; "
              f"threads {json.dumps(thread_report())}")
    kw = mode_kwargs(args.mode, N)
    line = {
        "impl": "reference", "metric": "isomp steps/sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(N, args.mode, kw, its),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample,
                         "threads": thread_report()},
        "e2e": {"value": val, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(N, mode, kw, its=None):
    cfg = {"workload": f"isomp on R({N},42): random skew-Hermitian trace-free complex128, norm_L2=1; "
                       f"{mode} mode dt={kw['dt'] / hbar(N):.2f}*hbar tol=auto maxit={kw['maxit']} minit={kw['minit']}",
           "N": N, "mode": mode,
           "l2": "working set (7 N^2 complex matrices + factor tables, > 500 MB at N=2048) exceeds the 126 MB L2; no flush needed"
           if N >= 2048 else "working set partly L2-resident as in production use; see DESIGN.md"}
    if its is not None:
        cfg["iterations_per_step"] = its
    return cfg


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import quflow_b200 as qf
    from quflow_b200._cuda import Handle, get_handle

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (quflow_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    N, mode = args.n, args.mode
    kw = mode_kwargs(mode, N)
    W0 = workload(N)
    handle = Handle(N, 1, local_rank) if world > 1 else get_handle(N, 1, local_rank)
    if world > 1:
        from quflow_b200.distributed import attach_row_sharding
        attach_row_sharding(handle, dist)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident -------------------------------------------------------------------
    W = torch.from_numpy(W0).to(dev)
    if args.warmup > 0:
        handle.isomp(W, kw["dt"], args.warmup, maxit=kw["maxit"], minit=kw["minit"])
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = handle.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res, _ = handle.isomp(W, kw["dt"], args.steps, maxit=kw["maxit"], minit=kw["minit"])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = handle.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    its = res[0]["total_iterations"] / max(args.steps, 1)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = args.steps / (ms * 1e-3)
    # run-to-run spread: the same K-step region again (the state simply keeps evolving), NOT part of `value`
    repeat_values = []
    for _ in range(max(args.repeats - 1, 0)):
        barrier()
        e0.record()
        handle.isomp(W, kw["dt"], args.steps, maxit=kw["maxit"], minit=kw["minit"])
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        repeat_values.append(args.steps / (float(t.item()) * 1e-3))

    # ---- e2e: public API, host buffers, one call per step ------------------------------------------
    Wpin = torch.from_numpy(W0.copy()).pin_memory()
    Wh = Wpin.numpy()
    # one GPU: the module-level qf.isomp.  Several GPUs: the host-buffer call on the row-sharded handle with a
    # row-distributed host state (tile-exchange path): every rank uploads its own two row blocks over its own PCIe link,
    # the state is completed over NVLink, the step runs sharded, every rank downloads its own rows again.  (Pull / NCCL
    # paths: every rank copies the whole state in and out.)
    own_rows = world > 1 and handle.comm_mode() == "tile"

    e2e_its = []

    def e2e_step():
        if world == 1:
            st = {"iterations": 0.0}
            qf.isomp(Wh, kw["dt"], steps=1, maxit=kw["maxit"], minit=kw["minit"], stats=st)
            e2e_its.append(st["iterations"])
        else:
            st, _ = handle.isomp(Wh, kw["dt"], 1, maxit=kw["maxit"], minit=kw["minit"], host_rows="own" if own_rows else "all")
            e2e_its.append(st[0]["total_iterations"])
    for _ in range(min(args.warmup, 2)):
        e2e_step()
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = e2e_steps / float(t.item())

    # what bounds e2e: the pinned-memory copies of one call, timed alone with CUDA events (the bytes this rank moves per
    # step, each direction), next to the device time of the step itself
    nb_rank = 16 * N * N // (world if own_rows else 1)
    stage = torch.empty(nb_rank, dtype=torch.uint8, device=dev)
    pin8 = Wpin.view(torch.uint8).reshape(-1)[:nb_rank]
    cp = {}
    for name, (dst, src) in (("h2d", (stage, pin8)), ("d2h", (pin8, stage))):
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            dst.copy_(src, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        cp[name] = c0.elapsed_time(c1) / 5.0
    del stage
    e2e_copy = {"bytes_per_rank_each_way": nb_rank, "h2d_ms": cp["h2d"], "d2h_ms": cp["d2h"],
                "h2d_GBps": nb_rank / cp["h2d"] / 1e6, "d2h_GBps": nb_rank / cp["d2h"] / 1e6,
                "e2e_ms_per_step": 1e3 / e2e_val, "device_ms_per_step": ms / max(args.steps, 1),
                "e2e_iterations_per_step": float(np.mean(e2e_its[-e2e_steps:])), "device_iterations_per_step": its,
                "note": "rank 0's pinned-memory copies of one call timed alone (CUDA events, 5 repeats).  e2e_ms_per_step - "
                        "device_ms_per_step = these two copies + the extra fixed-point iterations of a one-step call (every call "
                        "starts from dW = 0 like the reference, isospectral.py:430, so it cannot reuse the previous step's "
                        "increment as its first guess) + per-call set-up (tolerance norm, statistics read-back)"}

    ph_sharded = None
    if world > 1:
        # collective: per-phase device times of the SHARDED iteration (eager launches, gathers not overlapped)
        ph_sharded = handle.profile_iteration(torch.from_numpy(W0).to(dev), kw["dt"], reps=5)
    # ---- parity of the timed configuration on several GPUs: the sharded handle against a solo handle -----
    parity = None
    if world > 1:
        import hashlib
        psteps = 10
        Wp = torch.from_numpy(W0).to(dev)
        rsh, its_sh = handle.isomp(Wp, kw["dt"], psteps, maxit=kw["maxit"], minit=kw["minit"], want_iters=True)
        Wp_host = Wp.cpu().numpy()
        digests = [None] * world
        dist.all_gather_object(digests, hashlib.sha256(Wp_host.tobytes()).hexdigest())
        if rank == 0:
            hs = Handle(N, 1, local_rank)
            Ws = torch.from_numpy(W0).to(dev)
            rso, its_so = hs.isomp(Ws, kw["dt"], psteps, maxit=kw["maxit"], minit=kw["minit"], want_iters=True)
            Ws_host = Ws.cpu().numpy()
            hs.close()
            parity = {"against": f"a single-GPU handle on rank 0, {psteps} steps of the same R({N},42) workload",
                      "steps": psteps,
                      "rel_err": float(np.linalg.norm(Wp_host - Ws_host) / np.linalg.norm(Ws_host)),
                      "iterations_equal": bool(list(its_sh[0]) == list(its_so[0])),
                      "iterations": [int(x) for x in its_sh[0]],
                      "ranks_bit_identical": bool(all(d == digests[0] for d in digests))}
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0, single-GPU kernels) -------------------------------
    if world == 1:
        hsolo = handle
    else:
        from quflow_b200._cuda import Handle
        hsolo = Handle(N, 1, local_rank)      # a fresh, unsharded handle: single-GPU kernels only
    ph = hsolo.profile_iteration(torch.from_numpy(W0).to(dev), kw["dt"], reps=5)
    is3m, flops1, flops2 = hsolo.gemm_info()
    gemm1_tf = flops1 / (ph["gemm1_ms"] * 1e-3) / 1e12
    gemm2_tf = flops2 / (ph["gemm2_ms"] * 1e-3) / 1e12
    kname = ("k_zgemm_sk: A = P~ W~, %s complex arithmetic, %.4g executed real FP64 flop per launch "
             "(algorithmic 8 N^3 = %.4g)") % ("3M" if is3m else "4M", flops1, 8.0 * N ** 3)
    peaks = measured_peaks()
    hbm = peaks["hbm_gbs"] if peaks else 6650.0
    from quflow_b200._cuda.binding import measure_fp64_tensor_peak
    fp64_peak = measure_fp64_tensor_peak(local_rank, reps=5)
    pois_gbs = 32.0 * N * N / (ph["poisson_ms"] * 1e-3) / 1e9
    iter_ms = ph["poisson_ms"] + ph["gemm1_ms"] + ph["gemm2_ms"] + ph["post_ms"]
    roofline = {
        "bound": "tensor", "kernel": kname,
        "achieved": gemm1_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": gemm1_tf / fp64_peak,
        "traffic": ncu_traffic("k_zgemm3m_ws") if (N == 2048 and is3m) else None,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full capture at N=2048 "
                        "(profiles/r02_ncu_full_kernels.json); algorithmic operand bytes 3*16*N^2",
        "peak_source": "FP64 DMMA (mma.sync.m8n8k4.f64) issue peak measured on this GPU in this run "
                       "(qf_measure_fp64_tensor_peak, 16 warps/SM x 8 chains, best of 5); MEASURED_PEAKS.json has no FP64 "
                       f"entry; round-1 value {FP64_TENSOR_PEAK_TFLOPS_R01} TF/s, datasheet ~37-40 TF/s",
        "launch_ms": ph["gemm1_ms"],
        "second_gemm": {"achieved": gemm2_tf, "frac": gemm2_tf / fp64_peak, "launch_ms": ph["gemm2_ms"],
                        "executed_flop": flops2, "note": "S = A P~ is skew-Hermitian: lower-triangle tiles skipped"},
        "executed_flop": flops1, "algorithmic_tflops_equiv": 8.0 * N ** 3 / (ph["gemm1_ms"] * 1e-3) / 1e12,
        "share_of_iteration": (ph["gemm1_ms"] + ph["gemm2_ms"]) / iter_ms,
    }
    roofline_poisson = {
        "bound": "hbm", "kernel": "k_poisson_band: P~ = eps*Laplace^-1 W~ (32 N^2 algorithmic bytes: read W~, write P~)", "achieved": pois_gbs,
        "peak": hbm, "unit": "GB/s", "frac": pois_gbs / hbm,
        "traffic": ncu_traffic("k_poisson_band") if N == 2048 else None, "launch_ms": ph["poisson_ms"],
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "B200_PROFILING.md fallback 6.65 TB/s (of fallback)",
    }

    # ---- CPU baseline: bounded sample of the same workload on the host cores --------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = cpu_threads()
        force_cpu_threads(cores)
        n_cpu = 20 if N >= 2048 else (60 if N >= 1024 else 200)       # about 10-20 s of CPU work
        keep = {}
        cval, secs, cits, ckind, cwhat = run_cpu(N, mode, n_cpu, 1, keep)
        cpu = {"value": cval, "unit": "steps/s", "cores": cores, "kind": ckind, "threads": thread_report(),
               "sample": f"{n_cpu} steps (+1 warm-up) of the same R({N},42) workload, {cits:.2f} it/step, {secs:.1f} s; {cwhat}"}
        # parity of the timed configuration: the same two calls (1 step, then n_cpu steps) on the GPU against the CPU result
        Wg = torch.from_numpy(W0).to(dev)
        handle.isomp(Wg, kw["dt"], 1, maxit=kw["maxit"], minit=kw["minit"])
        rg, _ = handle.isomp(Wg, kw["dt"], n_cpu, maxit=kw["maxit"], minit=kw["minit"])
        Wg_host = Wg.cpu().numpy()
        parity = {"against": f"the CPU arm ({ckind}) on the same R({N},42) workload: 1 + {n_cpu} steps in two calls",
                  "steps": n_cpu + 1,
                  "rel_err": float(np.linalg.norm(Wg_host - keep["W"]) / np.linalg.norm(keep["W"])),
                  "iterations_equal": bool(rg[0]["total_iterations"] / n_cpu == cits),
                  "iterations_per_step": [rg[0]["total_iterations"] / n_cpu, cits]}

    line = {
        "metric": "isomp steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(N, mode, kw, its),
        "parallelism": ("single GPU" if world == 1 else
                        f"{world} GPUs: one simulation sharded by row blocks, data path '{handle.comm_mode()}' over NVLink "
                        f"peer memory (DESIGN.md section 4)"),
        "iterations_per_sec": value * its,
        "repeat_values": {"values": repeat_values, "note": f"the same {args.steps}-step region timed {len(repeat_values)} more "
                          "time(s) right after `value` (the state keeps evolving): run-to-run spread on this box"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "steps/s",
                "h2d_bytes_per_step": 16 * N * N * (1 if (world == 1 or own_rows) else world),
                "d2h_bytes_per_step": 16 * N * N * (1 if (world == 1 or own_rows) else world),
                "note": "one qf.isomp(W_numpy_pinned, dt, steps=1) call per step; on several GPUs the host state is "
                        "row-distributed (every rank copies its own 1/G of the rows in and out, bytes are the sum over ranks) "
                        "and completed over NVLink; chunked calls reset the warm start like the reference (isospectral.py:430)",
                "copies": e2e_copy},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_poisson": roofline_poisson,
        "phase_ms": ph,
        "phase_ms_sharded": ph_sharded,
        "incumbent": incumbents(N, dev) if world == 1 else None,
        "cpu_baseline": cpu,
        "parity": parity,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ensemble_arm(args):
    """BASELINE config 5: independent N=256 members, `--members` per GPU, sharded per member with no data-path
    collective (weak scaling).  value = member-steps/s over all ranks.  Every member has its own tolerance and its own
    convergence test (equivalent to separate reference calls); parity: two members per rank against the CPU oracle."""
    import torch
    import quflow_b200 as qf
    from quflow_b200._cuda import get_handle
    from quflow_b200._cuda.binding import measure_fp64_tensor_peak
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    N = 256 if args.n == 2048 else args.n
    k = args.members
    kw = mode_kwargs(args.mode, N)
    seeds = [1000 * rank + j for j in range(k)]
    W0 = np.stack([workload(N, seed=sd) for sd in seeds])
    handle = get_handle(N, k, local_rank)
    W = torch.from_numpy(W0).to(dev)
    handle.isomp(W, kw["dt"], max(args.warmup, 1), maxit=kw["maxit"], minit=kw["minit"])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = handle.launch_count()
    e0.record()
    res, _ = handle.isomp(W, kw["dt"], args.steps, maxit=kw["maxit"], minit=kw["minit"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = handle.launch_count() - l0
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # e2e: host buffers through the public API, one call per step
    Wpin = torch.from_numpy(W0.copy()).pin_memory()
    Wh = Wpin.numpy()
    qf.isomp_ensemble(Wh, kw["dt"], steps=1)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        qf.isomp_ensemble(Wh, kw["dt"], steps=1)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    # parity: the first and the last member of this rank against the CPU oracle, each run on its own like a reference call
    import oracle
    psteps = 20
    Wp = W0.copy()
    _, its = qf.isomp_ensemble(Wp, kw["dt"], steps=psteps, maxit=kw["maxit"], minit=kw["minit"], return_iterations=True)
    perr, pits = 0.0, True
    for j in sorted({0, k - 1}):
        rec = {}
        Wref = oracle.isomp(W0[j].copy(), kw["dt"], psteps, maxit=kw["maxit"], minit=kw["minit"], record=rec)
        perr = max(perr, float(np.linalg.norm(Wp[j] - Wref) / np.linalg.norm(Wref)))
        pits = pits and list(its[j]) == rec["iterations"]
    tp = torch.tensor([perr, 0.0 if pits else 1.0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    ph = handle.profile_iteration(torch.from_numpy(W0).to(dev), kw["dt"], reps=5)
    if rank == 0:
        is3m, flops1, flops2 = handle.gemm_info()
        peak = measure_fp64_tensor_peak(local_rank, reps=5)
        its_mean = float(np.mean([r["total_iterations"] for r in res])) / max(args.steps, 1)
        line = {"metric": "isomp member-steps/sec (ensemble)", "value": world * k * args.steps / (ms * 1e-3), "unit": "member-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
                "data": "synthetic",
                "config": {"workload": f"ensemble of {world * k} independent R({N},seed) members, {k} per GPU, {args.mode} mode, "
                                       f"each with its own tolerance and convergence (BASELINE config 5)", "N": N,
                           "members_per_gpu": k, "iterations_per_step": its_mean,
                           "l2": "working set of a rank (7 matrices x members, < 126 MB) is L2-resident as in production use"},
                "parallelism": f"{world} GPUs, members sharded per rank, no data-path collective",
                "clocks": clocks,
                "e2e": {"value": world * k * args.steps / e2e_s, "unit": "member-steps/s",
                        "h2d_bytes_per_step": 16 * N * N * k * world, "d2h_bytes_per_step": 16 * N * N * k * world,
                        "note": "one qf.isomp_ensemble(W_numpy_pinned[k,N,N], dt, steps=1) call per step on every rank"},
                "gpu_launches": launches,
                "roofline": {"bound": "tensor", "kernel": "k_zgemm3m_ws over the tile list of all members (A = P~ W~), executed flops",
                             "achieved": flops1 / (ph["gemm1_ms"] * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                             "frac": flops1 / (ph["gemm1_ms"] * 1e-3) / 1e12 / peak, "traffic": None, "launch_ms": ph["gemm1_ms"],
                             "executed_flop": flops1, "peak_source": "FP64 DMMA issue peak measured in this run"},
                "phase_ms": ph,
                "parity": {"against": f"the CPU oracle, members 0 and {k - 1} of every rank run on their own, {psteps} steps",
                           "rel_err": float(tp[0].item()), "iterations_equal": bool(tp[1].item() == 0.0)}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", "--size", dest="n", type=int, default=2048,
                    help="matrix size N (use --size under torch.distributed.run, whose own parser claims --n*)")
    ap.add_argument("--mode", default="natural", choices=["natural", "profile"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--repeats", type=int, default=3, help="timed regions in all: `value` is the first, the others are reported as spread")
    ap.add_argument("--workload", default="single", choices=["single", "ensemble"],
                    help="single: one R(N,42) simulation (row-sharded across GPUs); ensemble: BASELINE config 5, "
                         "--members independent N=256 simulations per GPU (weak scaling, no collective)")
    ap.add_argument("--members", type=int, default=8)
    args = ap.parse_args()
    if args.workload == "ensemble":
        ensemble_arm(args)
    elif args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
