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


def owner_of_row(i: int, N: int, world: int) -> int:
    """Rank that owns logical row i — and, on the tile-exchange path, every tile pair {(I, J), (J, I)}, I <= J, whose
    row block I contains i.  Mirrors qf_owner_of_row() in csrc/qf_common.cuh."""
    if world == 1:
        return 0
    blk = i // (N // (2 * world))
    return blk if blk < world else 2 * world - 1 - blk


def tile_exchange_supported(N: int, world: int) -> bool:
    """The tile-exchange data path needs ownership blocks made of whole 64-row tiles: N divisible by 128 * world
    (csrc/comm.cu: p2p_finish); otherwise the library falls back to the pull all-gather."""
    return world == 1 or N % (128 * world) == 0


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


class ShardedIsomp:
    """``isomp`` over the ranks of a torch.distributed process group (one process per GPU) — the multi-GPU counterpart of
    the reference's pluggable integrator object (``IsompCUDA``, quflow/experimental/isospectral_cuda.py:120-358, passed
    as ``integrator=`` to ``qf.solve``, quflow/simulation.py:554-566).

    Every rank calls it with the same arguments and the same (replicated) state; the state that comes back is complete
    and bit-identical on every rank.  ``quflow_b200.solve(W, ..., integrator=ShardedIsomp(dist))`` keeps the state on the
    GPUs between output intervals; give the output callback (``QuSimulation``) to rank 0 only.

    The parameter list mirrors ``isomp_fixedpoint`` (isospectral.py:338-353) so that ``solve`` finds ``stats`` by
    introspection (simulation.py:729).  Hooks that run host code inside the step (``callback``, ``forcing``,
    ``strang_splitting``, custom Hamiltonians) run on EVERY rank; that mode shards the two GEMMs and completes their outputs
    on every rank (all-gather path), so the hooks see complete, identical matrices.
    """
    device_resident = True      # solve(): keep the state on the device between output intervals

    def __init__(self, dist, device=None, mode=None):
        import torch
        self.dist = dist
        self.mode = mode
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self._handles = {}

    def handle(self, N):
        if N not in self._handles:
            h = binding.Handle(N, 1, self.device.index)
            attach_row_sharding(h, self.dist, mode=self.mode)
            self._handles[N] = h
        return self._handles[N]

    def __call__(self, W, dt, steps=100, hamiltonian=None, time=None, forcing=None, strang_splitting=None, stats=None,
                 callback=None, tol='auto', maxit=10, minit=1, verbatim=False, compsum=False, reinitialize=False):
        from .integrators import _is_default_hamiltonian
        from .laplacian import _is_torch
        assert minit >= 1, "minit must be at least 1."          # isospectral.py:400
        assert maxit >= minit, "maxit must be at minit."         # isospectral.py:401
        if W.ndim != 2:
            raise NotImplementedError("ShardedIsomp advances one (N, N) state; ensembles shard per member (member_slice)")
        if forcing is not None or strang_splitting is not None or callback is not None or not _is_default_hamiltonian(hamiltonian):
            # hooks run host code on every rank: host-stepped mode, GEMMs sharded, everything else replicated; the hooks
            # receive arrays of the same kind as W (numpy copies or device tensors), as on one GPU
            from .integrators import _isomp_host_stepped
            return _isomp_host_stepped(W, dt, steps, hamiltonian, time, forcing, strang_splitting, stats, callback, tol, maxit,
                                       minit, verbatim and self.dist.get_rank() == 0, compsum, reinitialize,
                                       handle=self.handle(W.shape[-1]))
        auto = (isinstance(tol, str) and tol == 'auto') or (not isinstance(tol, str) and tol < 0)
        if _is_torch(W):
            Wc = W if W.is_contiguous() else W.contiguous()
        else:
            Wc = np.ascontiguousarray(W, dtype=np.complex128)
        res, _ = self.handle(Wc.shape[-1]).isomp(Wc, dt, steps, tol=-1.0 if auto else float(tol), maxit=maxit, minit=minit,
                                                  compsum=bool(compsum), reinitialize=bool(reinitialize))
        if Wc is not W:
            if _is_torch(W):
                W.copy_(Wc)
            else:
                W[...] = Wc
        st = res[0]
        if auto and stats:
            stats['tol_auto'] = st['tol_used']                   # :451-452
        if verbatim and steps > 0 and self.dist.get_rank() == 0:
            print("Average number of iterations per step: {:.2f}".format(st['total_iterations'] / steps))
        if stats and steps > 0:                                   # :609-611
            stats["iterations"] = st['total_iterations'] / steps
            stats["number_of_maxit"] = st['number_of_maxit'] / steps
        return W

    def close(self):
        for h in self._handles.values():
            h.close()
        self._handles = {}
