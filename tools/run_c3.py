#!/usr/bin/env python
"""BASELINE config 3: N=1024 isomp long run with output through QuSimulation, 1 vs 2 GPUs.

    python tools/run_c3.py [--steps 3000] [--steps-out 100]                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/run_c3.py

`quflow_b200.solve` drives the run exactly like the reference's `qf.solve` (quflow/simulation.py:584-802): R(1024, 42),
dt = 0.25*hbar, one output record every `steps_out` steps appended by a QuSimulation callback on rank 0 (reference layout:
`mat` (T, N, N) chunked (1, N, N), `time`, `step`, statistics series).  The state stays on the GPU(s); every record is one
asynchronous D2H copy that overlaps the next chunk.  h5py / libhdf5 are not part of this image, so the HDF5 sink is the
in-memory stand-in of the test-suite (tests/fake_h5py.py) unless a real h5py is importable; the line says which.
Prints one JSON line: steps/s end to end (host array in, records out), and the same run without output for reference.
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--steps-out", type=int, default=100)
    args = ap.parse_args()
    import torch
    import quflow_b200 as qf
    from bench import workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    integrator = qf.isomp
    dist = None
    if world > 1:
        import torch.distributed as dist
        from quflow_b200.distributed import ShardedIsomp
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        integrator = ShardedIsomp(dist)
    try:
        import h5py  # noqa: F401
        sink = "h5py"
    except ImportError:
        import fake_h5py
        sys.modules["h5py"] = fake_h5py
        sink = "in-memory stand-in for h5py (tests/fake_h5py.py; h5py is not installed in this image)"
    N = args.size
    W0 = workload(N)
    dt = 0.25 * qf.hbar(N)

    def run(with_output):
        W = W0.copy()
        cb = None
        tmp = None
        if with_output and rank == 0:
            tmp = os.path.join(tempfile.mkdtemp(), "c3.hdf5")
            cb = qf.QuSimulation(tmp, state=W, overwrite=True)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        qf.solve(W, dt=dt, steps=args.steps, steps_out=args.steps_out, integrator=integrator, callback=cb, progress_bar=False)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        secs = time.perf_counter() - t0
        records = None
        if cb is not None:
            records = int(cb['mat'].shape[0])
        return secs, W, records

    run(False)                                   # warm-up: handles, graphs, pinned buffers
    s_plain, W_plain, _ = run(False)
    s_out, W_out, records = run(True)
    if rank == 0:
        line = {"metric": "isomp steps/sec end to end through solve() with QuSimulation output (BASELINE config 3)",
                "value": args.steps / s_out, "unit": "steps/s", "n_gpus": world, "steps": args.steps, "steps_out": args.steps_out,
                "N": N, "records_written": records, "bytes_per_record": 16 * N * N, "sink": sink,
                "value_without_output": args.steps / s_plain,
                "output_overhead_frac": s_out / s_plain - 1.0,
                "same_state_with_and_without_output": bool(np.array_equal(W_plain, W_out))}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
