"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: partition maps and the unique-id exchange."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quflow_b200 import distributed as qd


@pytest.mark.parametrize("N,world", [(1024, 2), (2048, 8), (2048, 4), (512, 1), (96, 3)])
def test_row_partition_is_a_balanced_permutation(N, world):
    blocks = qd.row_blocks(N, world)
    rows = sorted(i for rb in blocks for (a, b) in rb for i in range(a, b))
    assert rows == list(range(N))                                   # every row owned exactly once
    assert len({sum(b - a for a, b in rb) for rb in blocks}) == 1   # equal row counts (NCCL all-gather needs it)
    perm = [qd.permuted_row(i, N, world) for i in range(N)]
    assert sorted(perm) == list(range(N))
    if world > 1:
        hb = N // (2 * world)
        for r, rb in enumerate(blocks):
            got = sorted(qd.permuted_row(i, N, world) for (a, b) in rb for i in range(a, b))
            assert got == list(range(2 * r * hb, 2 * (r + 1) * hb))   # a rank's rows are contiguous after permutation
        # upper-triangular work (number of (i, j >= i) pairs) is balanced to within one block
        work = [sum(N - i for (a, b) in rb for i in range(a, b)) for rb in blocks]
        assert max(work) - min(work) <= hb * hb


@pytest.mark.parametrize("N,world", [(256, 2), (1024, 4), (2048, 8), (1024, 8)])
def test_tile_pair_ownership_is_a_balanced_partition(N, world):
    """Tile-exchange path: the upper 64 x 64 tiles (I <= J) belong to the owner of row block I.  Every tile has exactly one
    owner, the owners agree with row_blocks(), and pairing block r with block 2*world-1-r balances the upper-triangular
    work: tile counts per rank differ by at most one tile column."""
    assert qd.tile_exchange_supported(N, world)
    nt = N // 64
    counts = [0] * world
    for I in range(nt):
        owner = qd.owner_of_row(64 * I, N, world)
        assert all(qd.owner_of_row(r, N, world) == owner for r in range(64 * I, 64 * I + 64))      # whole tiles
        assert any(a <= 64 * I < b for a, b in qd.row_blocks(N, world)[owner])
        counts[owner] += nt - I                                                                    # tiles (I, J >= I)
    assert sum(counts) == nt * (nt + 1) // 2
    assert max(counts) - min(counts) <= N // (2 * world) // 64
    assert not qd.tile_exchange_supported(N + 64, world)


def test_row_partition_rejects_bad_sizes():
    with pytest.raises(ValueError):
        qd.row_blocks(1000, 8)


@pytest.mark.parametrize("k,world", [(64, 8), (10, 4), (3, 8)])
def test_member_slices_cover_the_ensemble(k, world):
    seen = []
    for r in range(world):
        s = qd.member_slice(k, r, world)
        seen += list(range(k))[s]
    assert seen == list(range(k))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _FakeHandle:
    """Stands in for binding.Handle: records what attach_row_sharding asks the library to do (no CUDA)."""

    def __init__(self, N, rank):
        self.N, self.rank, self.calls = N, rank, []

    def p2p_export(self):
        return bytes([self.rank]) * 256

    def p2p_import(self, blobs, rank, world):
        self.calls.append(("import", [b[0] for b in blobs], rank, world))

    def comm_set_tile(self, enable):
        self.calls.append(("tile", bool(enable)))          # True: tile exchange, False: pull all-gather

    def comm_init(self, uid, rank, world):
        self.calls.append(("nccl", len(uid), rank, world))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fake = bytes(range(128))
        uid = qd.broadcast_unique_id(dist, make_id=lambda: fake)   # no CUDA needed: id creation is injected
        ok = uid == fake
        # ensemble sharding: every rank owns a disjoint slice; gathering the slices restores the order
        k = 5
        mine = torch.arange(k)[qd.member_slice(k, rank, world)]
        gathered = [None] * world
        dist.all_gather_object(gathered, mine.tolist())
        ok = ok and sum(gathered, []) == list(range(k))
        # row sharding set-up: every rank imports every rank's IPC blob in rank order; the data path follows `mode`
        h = _FakeHandle(1024, rank)
        qd.attach_row_sharding(h, dist, mode="p2p")
        ok = ok and h.calls == [("import", list(range(world)), rank, world)]          # default: the library picks tile/pull
        for mode, want in (("tile", True), ("pull", False)):
            h = _FakeHandle(1024, rank)
            qd.attach_row_sharding(h, dist, mode=mode)
            ok = ok and h.calls == [("import", list(range(world)), rank, world), ("tile", want)]
        try:
            qd.attach_row_sharding(_FakeHandle(1002, rank), dist)                    # N not divisible by 2 * world
            ok = False
        except ValueError:
            pass
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_unique_id_broadcast_and_ensemble_gather_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert dict(out) == {0: True, 1: True}
