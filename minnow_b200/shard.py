"""Block-range sharding of one minnow/minp file over the ranks of a process group
(SURVEY.md 8e).  Blocks are independent given the group parameters, so rank r encodes a
contiguous range of blocks into its own buffer with rank-LOCAL byte offsets; the only
exchange is an all-gather of the per-block packed sizes, after which every rank derives
the same global offsets (go/block_index.go:16-35 applied to the whole file) and knows
where its bytes go.  Payload never crosses the interconnect.

Works on CPU tensors with the gloo backend (tests) and on CUDA tensors with NCCL
(bench.py --gpus N); without a process group it degenerates to the single-rank scan."""
import torch
import torch.distributed as dist


def block_range(nblocks, rank, world):
    """Contiguous, balanced partition: rank r owns blocks [lo, hi)."""
    base, rem = divmod(int(nblocks), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def packed_sizes(bits, n):
    """bit.ArrayBytes(bits, n) per block (go/bit/bit.go:23-25), as int64 tensor."""
    return (bits.to(torch.int64) * int(n) + 7) >> 3


def exclusive_scan(sizes, base=0, ctx=None):
    """blockOffset of every block (go/block_index.go:25-35) + total.  On CUDA tensors the
    library's k_scan_sizes does the scan (ctx = minnow_b200.Context); CPU tensors (gloo
    tests) use cumsum."""
    sizes = sizes.to(torch.int64).contiguous()
    if sizes.is_cuda and ctx is not None:
        offs = torch.empty_like(sizes)
        total = torch.zeros(1, dtype=torch.int64, device=sizes.device)
        ctx.scan_offsets_dev(sizes, sizes.numel(), int(base), offs, total)
        return offs, total
    inc = torch.cumsum(sizes, 0)
    return inc - sizes + int(base), inc[-1:].clone() if sizes.numel() else torch.zeros(1, dtype=torch.int64)


def gather_sizes(local_sizes, counts=None, group=None):
    """All-gather of per-block sizes.  counts = blocks per rank (list) when ranks own
    different numbers of blocks; None when every rank owns local_sizes.numel() blocks."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local_sizes.clone()
    world = dist.get_world_size(group)
    if counts is None:
        out = torch.empty(world * local_sizes.numel(), dtype=local_sizes.dtype, device=local_sizes.device)
        dist.all_gather_into_tensor(out, local_sizes.contiguous(), group=group)
        return out
    width = max(counts)
    padded = torch.zeros(width, dtype=local_sizes.dtype, device=local_sizes.device)
    padded[:local_sizes.numel()] = local_sizes
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


def global_offsets(local_sizes, nblocks_total, group_blocks, group=None, ctx=None):
    """Global byte offsets of a file whose `nblocks_total` blocks are split by block_range.

    local_sizes : int64 [hi - lo] packed size of this rank's blocks
    group_blocks: blocks per minnow group (offsets restart at 0 in every group: each group has
                  its own data region, go/writer.go:84-86)
    -> (offsets of ALL blocks within their groups [nblocks_total], group sizes [ngroups],
        (lo, hi) of this rank)
    """
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    counts = [block_range(nblocks_total, r, world)[1] - block_range(nblocks_total, r, world)[0] for r in range(world)]
    assert local_sizes.numel() == counts[rank]
    sizes = gather_sizes(local_sizes, counts if len(set(counts)) > 1 else None, group)
    assert sizes.numel() == nblocks_total and nblocks_total % group_blocks == 0
    offs, _ = exclusive_scan(sizes, 0, ctx)                      # one scan over the whole file ...
    per_group = sizes.view(-1, group_blocks)
    starts = offs.view(-1, group_blocks)[:, :1]                  # ... re-based at every group start
    return (offs.view(-1, group_blocks) - starts).reshape(-1), per_group.sum(1), block_range(nblocks_total, rank, world)
