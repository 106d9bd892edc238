#!/usr/bin/env python
"""bench.py — isomp steps/s on B200 (BASELINE.json metric), roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--n 2048] [--mode natural|profile]

Workload (config.workload): R(N, 42) — random skew-Hermitian trace-free complex128 vorticity normalised to
norm_L2 = 1, N = 2048 by default (the size BASELINE.json's metric is quoted on; it fits one GPU),
natural mode: dt = 0.25*hbar(N), tol='auto', maxit=10, minit=1 (about 3 fixed-point iterations per step).
A "step" is one isospectral-midpoint time step.

* ``value``  : steps/s with W resident in HBM, timed with CUDA events over exactly K steps (max over ranks).
* ``e2e``    : the same metric through the public Python API with HOST buffers (numpy in pinned memory): every step
               is one ``qf.isomp(W_host, dt, steps=1)`` call = H2D copy of W + one step + D2H copy of W.  On several
               GPUs the same host-buffer call goes to the row-sharded handle: every rank copies the replicated state
               in over its own PCIe link, the step runs sharded, every rank reads the result back.
* ``roofline``: the dominant kernel (k_zgemm3m_ws, FP64 DMMA) — EXECUTED flops per launch / CUDA-event launch time,
               against the FP64 tensor peak measured on this pool (MEASURED_PEAKS.json has no FP64 entry; see
               profiles/r01_fp64_pipes.txt).  ``roofline_poisson`` reports the HBM-bound Poisson solve against
               MEASURED_PEAKS.json's copy bandwidth.
* ``cpu_baseline`` / ``--impl reference``: the CPU oracle port of the reference path (numpy BLAS zgemm + OpenMP
               Thomas, all host cores) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

FP64_TENSOR_PEAK_TFLOPS = 37.15   # measured DMMA m8n8k4 issue peak, 148 SMs @ 1965 MHz (profiles/r01_fp64_pipes.txt)


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
    (profiles/r01_ncu_full_kernels.json, N=2048); None for other sizes or when the file is missing."""
    try:
        rows = [r for r in json.load(open(os.path.join(ROOT, "profiles", "r01_ncu_full_kernels.json")))
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
def cpu_threads():
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    return n


def run_cpu_port(N, mode, steps, warmup):
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
    return steps / dt, dt, stats['iterations']


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.n
    cores = cpu_threads()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    # bounded: at N=2048 one CPU step costs seconds; keep the whole run within a few minutes
    warm = min(args.warmup, 1)
    val, secs, its = run_cpu_port(N, args.mode, args.steps, warm)
    sample = f"{args.steps} steps (+{warm} warm-up) of R({N},42), {args.mode} mode, {its:.2f} it/step, {secs:.1f} s"
    kw = mode_kwargs(args.mode, N)
    line = {
        "impl": "reference", "metric": "isomp steps/sec", "value": val, "unit": "steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(N, args.mode, kw, its),
        "cpu_baseline": {"value": val, "unit": "steps/s", "cores": cores, "kind": "port", "sample": sample},
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

    # ---- e2e: public API, host buffers, one call per step ------------------------------------------
    Wpin = torch.from_numpy(W0.copy()).pin_memory()
    Wh = Wpin.numpy()
    # one GPU: the module-level qf.isomp; several GPUs: the same host-buffer call on the row-sharded handle (every rank
    # copies the replicated state in over its own PCIe link, the step runs sharded, every rank reads the result back)
    def e2e_step():
        if world == 1:
            qf.isomp(Wh, kw["dt"], steps=1, maxit=kw["maxit"], minit=kw["minit"])
        else:
            handle.isomp(Wh, kw["dt"], 1, maxit=kw["maxit"], minit=kw["minit"])
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

    ph_sharded = None
    if world > 1:
        # collective: per-phase device times of the SHARDED iteration (eager launches, gathers not overlapped)
        ph_sharded = handle.profile_iteration(torch.from_numpy(W0).to(dev), kw["dt"], reps=5)
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
    pois_gbs = 32.0 * N * N / (ph["poisson_ms"] * 1e-3) / 1e9
    iter_ms = ph["poisson_ms"] + ph["gemm1_ms"] + ph["gemm2_ms"] + ph["post_ms"]
    roofline = {
        "bound": "tensor", "kernel": kname,
        "achieved": gemm1_tf, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": gemm1_tf / FP64_TENSOR_PEAK_TFLOPS,
        "traffic": ncu_traffic("k_zgemm3m_ws") if (N == 2048 and is3m) else None,
        "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of this launch, ncu --set full capture at N=2048 "
                        "(profiles/r01_ncu_full_kernels.json); algorithmic operand bytes 3*16*N^2",
        "peak_source": "measured FP64 DMMA issue peak on this pool's B200 (profiles/r01_fp64_pipes.txt); "
                       "MEASURED_PEAKS.json has no FP64 entry, datasheet ~37-40 TF/s",
        "launch_ms": ph["gemm1_ms"],
        "second_gemm": {"achieved": gemm2_tf, "frac": gemm2_tf / FP64_TENSOR_PEAK_TFLOPS, "launch_ms": ph["gemm2_ms"],
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
        os.environ.setdefault("OMP_NUM_THREADS", str(cores))
        n_cpu = 3 if N >= 2048 else (6 if N >= 1024 else 20)
        cval, secs, cits = run_cpu_port(N, mode, n_cpu, 1)
        cpu = {"value": cval, "unit": "steps/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} steps (+1 warm-up) of the same R({N},42) workload, {cits:.2f} it/step, {secs:.1f} s; "
                         f"numpy BLAS zgemm + OpenMP Thomas (oracle/)"}

    line = {
        "metric": "isomp steps/sec", "value": value, "unit": "steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 (complex128)", "data": "synthetic",
        "config": workload_config(N, mode, kw, its),
        "parallelism": ("single GPU" if world == 1 else
                        f"{world} GPUs: GEMMs sharded by row blocks, state replicated, {handle.comm_mode()} "
                        f"all-gather over NVLink peer memory (DESIGN.md section 4)"),
        "iterations_per_sec": value * its,
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": "steps/s", "h2d_bytes_per_step": 16 * N * N, "d2h_bytes_per_step": 16 * N * N,
                "note": "one qf.isomp(W_numpy_pinned, dt, steps=1) call per step (on several GPUs: the same host-buffer call "
                        "on the row-sharded handle, every rank copying in and out); chunked calls reset the warm start "
                        "like the reference (isospectral.py:430)"},
        "gpu_launches": launches,
        "roofline": roofline,
        "roofline_poisson": roofline_poisson,
        "phase_ms": ph,
        "phase_ms_sharded": ph_sharded,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def ensemble_arm(args):
    """BASELINE config 5: independent N=256 members, `--members` per GPU, sharded per member with no data-path
    collective (weak scaling).  value = member-steps/s over all ranks."""
    import torch
    import quflow_b200 as qf
    from quflow_b200._cuda import get_handle
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
    W0 = np.stack([workload(N, seed=1000 * rank + j) for j in range(k)])
    handle = get_handle(N, k, local_rank)
    W = torch.from_numpy(W0).to(dev)
    handle.isomp(W, kw["dt"], args.warmup, maxit=kw["maxit"], minit=kw["minit"])
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = handle.launch_count()
    e0.record()
    res, _ = handle.isomp(W, kw["dt"], args.steps, maxit=kw["maxit"], minit=kw["minit"])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = handle.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    # e2e: host buffers through the public API, one call per step
    Wpin = torch.from_numpy(W0.copy()).pin_memory()
    Wh = Wpin.numpy()
    qf.isomp_ensemble(Wh, kw["dt"], steps=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        qf.isomp_ensemble(Wh, kw["dt"], steps=1)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    if rank == 0:
        its = float(np.mean([r["total_iterations"] for r in res])) / max(args.steps, 1)
        line = {"metric": "isomp member-steps/sec (ensemble)", "value": world * k * args.steps / (ms * 1e-3), "unit": "member-steps/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / max(args.steps, 1),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64 (complex128)",
                "data": "synthetic",
                "config": {"workload": f"ensemble of {world * k} independent R({N},seed) members, {k} per GPU, {args.mode} mode, "
                                       f"each with its own tolerance and convergence", "N": N, "members_per_gpu": k,
                           "iterations_per_step": its},
                "e2e": {"value": world * k * args.steps / e2e_s, "unit": "member-steps/s",
                        "h2d_bytes_per_step": 16 * N * N * k, "d2h_bytes_per_step": 16 * N * N * k},
                "gpu_launches": launches}
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
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--mode", default="natural", choices=["natural", "profile"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
