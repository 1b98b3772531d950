"""World-size-2 (and 3) gloo tests of the sharded path's host logic (minnow_b200/shard.py):
block-range partition, all-gather of per-block packed sizes, global offset scan.  The
single-process result is the reference's own running sum (go/block_index.go:16-35)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from minnow_b200 import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nblocks, group_blocks, n, seed, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(seed)
        bits = torch.from_numpy(rng.integers(0, 33, nblocks))           # every rank derives the same file
        lo, hi = shard.block_range(nblocks, rank, world)
        local = shard.packed_sizes(bits[lo:hi], n)
        offs, gsizes, rng_ = shard.global_offsets(local, nblocks, group_blocks)
        q.put((rank, offs.numpy(), gsizes.numpy(), rng_))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,nblocks,group_blocks", [(2, 3 * 64, 64), (2, 3 * 9, 9), (3, 3 * 8, 8)])
def test_global_offsets_match_single_process_scan(world, nblocks, group_blocks):
    n, seed = 4096, 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nblocks, group_blocks, n, seed, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 33, nblocks)
    sizes = (bits * n + 7) >> 3
    want = np.concatenate([np.concatenate([[0], np.cumsum(g)[:-1]]) for g in sizes.reshape(-1, group_blocks)])
    covered = []
    for rank, offs, gsizes, (lo, hi) in res:
        assert np.array_equal(offs, want)                               # every rank holds the same global index
        assert np.array_equal(gsizes, sizes.reshape(-1, group_blocks).sum(1))
        covered += list(range(lo, hi))
    assert sorted(covered) == list(range(nblocks))                      # disjoint cover of the blocks


def test_block_range_is_balanced_and_contiguous():
    for nblocks in (0, 1, 7, 192, 98304):
        for world in (1, 2, 3, 8):
            r = [shard.block_range(nblocks, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == nblocks
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            lens = [b - a for a, b in r]
            assert max(lens) - min(lens) <= 1


def test_single_rank_degenerates_to_plain_scan():
    bits = torch.tensor([0, 10, 3, 17, 0, 1])
    sizes = shard.packed_sizes(bits, 100)
    offs, gsizes, (lo, hi) = shard.global_offsets(sizes, 6, 3)
    assert offs.tolist() == [0, 0, 125, 0, 213, 213] and gsizes.tolist() == [163, 226] and (lo, hi) == (0, 6)
