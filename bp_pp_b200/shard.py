"""Sharding of a proof batch across the GPUs of one box: one process per GPU, contiguous slices, no
data-path collective (proofs are independent -- the reference has no shared state between calls,
src/range_proof/u64_proof.rs:42-82).  Generators/tables are replicated per GPU by each rank's Context.
`gather_status` is a convenience for callers that want the whole verdict vector on every rank."""
from __future__ import annotations

from typing import List, Sequence, Tuple


def shard_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of rank `rank`; sizes differ by at most one; concatenation over ranks is 0..n."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_bytes(buf: bytes, item: int, n: int, world: int, rank: int) -> bytes:
    lo, hi = shard_bounds(n, world, rank)
    return buf[item * lo:item * hi]


def gather_status(local: Sequence[int], n: int) -> List[int]:
    """All ranks contribute their slice of verdicts; returns the full vector (needs torch.distributed)."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_bounds(n, world, r)[1] - shard_bounds(n, world, r)[0] for r in range(world)]
    assert len(local) == sizes[rank]
    mx = max(sizes) if sizes else 0
    t = torch.full((mx,), -99, dtype=torch.int32)
    t[:len(local)] = torch.tensor(list(local), dtype=torch.int32)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full: List[int] = []
    for r in range(world):
        full += outs[r][:sizes[r]].tolist()
    return full


def msm_sharded(points: bytes, scalars32: bytes, device: int, points_fmt: int = 1) -> bytes:
    """One large MSM split by point range over the ranks of the process group (one GPU each): every rank computes the
    partial sum of its contiguous block on its own GPU, the 33-byte partial points are all-gathered (NCCL when the group
    is NCCL: the only data-path collective in this package, a few hundred bytes) and every rank adds them.
    `points`/`scalars32` hold the FULL vectors on every rank (generators are replicated); returns the 33-byte sum."""
    import torch
    import torch.distributed as dist
    from . import api
    world, rank = dist.get_world_size(), dist.get_rank()
    psz = 33 if points_fmt == api.FMT_COMPRESSED else 64
    n = min(len(points) // psz, len(scalars32) // 32)
    lo, hi = shard_bounds(n, world, rank)
    part = api.msm(points[psz * lo:psz * hi], scalars32[32 * lo:32 * hi], points_fmt, api.FMT_COMPRESSED, device)
    use_cuda = dist.get_backend() == "nccl"
    dev = torch.device("cuda", device) if use_cuda else torch.device("cpu")
    mine = torch.frombuffer(bytearray(part), dtype=torch.uint8).to(dev)
    gathered = [torch.empty(33, dtype=torch.uint8, device=dev) for _ in range(world)]
    dist.all_gather(gathered, mine)
    allparts = b"".join(bytes(t.cpu().numpy().tobytes()) for t in gathered)
    return api.points_sum(allparts, api.FMT_COMPRESSED, api.FMT_COMPRESSED, device)
