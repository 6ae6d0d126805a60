"""Proof-parallel sharding across ranks: independent proofs are
dealt to ranks in contiguous blocks, the CRS is replicated, there is no
data-path collective; only the verdict bitmap is gathered."""
from __future__ import annotations


def shard_bounds(total: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `total` units owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad rank/world")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def gather_verdicts(local_ok, total: int, world: int, rank: int, device="cpu"):
    """All ranks obtain the full verdict list (length `total`) from the per-rank slices.
    Works with any torch.distributed backend (nccl on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist

    lo, hi = shard_bounds(total, world, rank)
    if len(local_ok) != hi - lo:
        raise ValueError("local verdict slice has the wrong length")
    if world == 1:
        return [int(v) for v in local_ok]
    width = -(-total // world)
    buf = torch.full((width,), -1, dtype=torch.int32, device=device)
    if hi > lo:
        buf[: hi - lo] = torch.tensor([int(v) for v in local_ok], dtype=torch.int32, device=device)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    out = []
    for r in range(world):
        a, b = shard_bounds(total, world, r)
        out.extend(int(v) for v in parts[r][: b - a].tolist())
    return out
