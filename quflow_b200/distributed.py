"""Multi-GPU host logic: one process per GPU, torch.distributed for the plumbing, NCCL inside the library.

New functionality relative to the reference (which is single-process):

* ``attach_row_sharding(handle, dist)`` — shard ONE large-N simulation by row blocks across the ranks of the default
  process group.  All per-iteration communication runs inside libquflow_b200.so, on the compute stream, with no Python
  in the loop (peer-memory kernels, or NCCL); torch.distributed only carries the set-up blobs (CUDA IPC handles or the
  128-byte ncclUniqueId).
* ``member_slice(k, rank, world)`` / ``isomp_ensemble_sharded`` — shard an ensemble per member: no data-path
  collective at all, one gather of the results at the end.
"""
import numpy as np

from ._cuda import binding


def row_blocks(N: int, world: int):
    """Logical row ranges owned by each rank: blocks r and 2*world-1-r of N/(2*world) rows (balanced for the
    upper-triangular second GEMM).  Mirrors qf_prow()/qf_block_rows() in csrc/qf_common.cuh."""
    if world == 1:
        return [[(0, N)]]
    if N % (2 * world) != 0:
        raise ValueError(f"row sharding needs N divisible by 2*world (N={N}, world={world})")
    hb = N // (2 * world)
    return [[(r * hb, (r + 1) * hb), ((2 * world - 1 - r) * hb, (2 * world - r) * hb)] for r in range(world)]


def permuted_row(i: int, N: int, world: int) -> int:
    """Row index of logical row i in the rank-permuted layout of the GEMM outputs."""
    if world == 1:
        return i
    hb = N // (2 * world)
    blk = i // hb
    r, slot = (blk, 0) if blk < world else (2 * world - 1 - blk, 1)
    return (2 * r + slot) * hb + (i - blk * hb)


def broadcast_unique_id(dist, make_id=None) -> bytes:
    """Rank 0 creates the ncclUniqueId; everybody receives it through the default process group (any backend)."""
    import torch
    rank = dist.get_rank()
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(binding.QF_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = (make_id or binding.comm_unique_id)()
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def attach_row_sharding(handle, dist, mode=None):
    """Make ``handle`` (batch == 1) run one simulation row-sharded over the default process group.

    mode "p2p" (default): peer-memory data path chosen by the library — "tile" whenever N is divisible by 128*world,
    else "pull" (csrc/comm.cu).  torch.distributed only all-gathers one 256-byte blob of CUDA IPC handles per rank at
    set-up; everything per iteration runs inside the step graph as kernels over the NVLink peer mappings.
    mode "tile": tile exchange — GEMMs, the tail of the iteration and the update are all sharded; per iteration a rank
    pushes the lower tiles of A its peers need (from the GEMM epilogue) and its new W~ tiles; Poisson runs replicated.
    mode "pull": the GEMM outputs A and S are completed on every rank by pull kernels; tail and update run replicated.
    mode "nccl": one in-place ncclAllGather per GEMM, issued by the library on the compute stream (eager launches;
    NCCL cannot run inside the conditional graph body).  Select with QF_COMM=tile|pull|nccl.
    """
    import os
    world, rank = dist.get_world_size(), dist.get_rank()
    if world == 1:
        return handle
    row_blocks(handle.N, world)     # validates divisibility with a Python-level error
    mode = mode or os.environ.get("QF_COMM", "p2p")
    if mode == "nccl":
        uid = broadcast_unique_id(dist)
        handle.comm_init(uid, rank, world)
    else:
        blobs = [None] * world
        dist.all_gather_object(blobs, handle.p2p_export())
        handle.p2p_import(blobs, rank, world)       # picks tile or pull (csrc/comm.cu)
        if mode in ("tile", "pull"):
            handle.comm_set_tile(mode == "tile")
        dist.barrier()              # every rank has mapped every peer before anybody starts signalling
    return handle


def member_slice(k: int, rank: int, world: int) -> slice:
    """Members of a k-member ensemble owned by ``rank`` (contiguous, sizes differ by at most one)."""
    base, rem = divmod(k, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def isomp_ensemble_sharded(W_local, dt, steps, **kw):
    """Advance this rank's members of an ensemble (``W_local``: (k_local, N, N)); no collective on the data path."""
    from .integrators import isomp_ensemble
    return isomp_ensemble(W_local, dt, steps, **kw)
