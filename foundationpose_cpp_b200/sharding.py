"""Host-side arithmetic of hypothesis sharding (mirrors fpb200_session::set_shard and
unpack_exchange_kernel in csrc/session.cu).  The reference has no multi-GPU path; hypotheses are
independent up to the scorer's cross-hypothesis attention, so each rank takes a contiguous range and
the ranks exchange one [528]-float row (512-d feature + refined pose) per hypothesis."""
from __future__ import annotations

from typing import Tuple

EXCH_LD = 512 + 16


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of rank `rank`: the first n % world ranks get one extra hypothesis
    (252 over 8 ranks -> 32,32,32,32,31,31,31,31)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_capacity(n: int, world: int) -> int:
    """Rows per rank in the padded all-gather buffer."""
    return (n + world - 1) // world


def padded_row(n: int, world: int, row: int) -> int:
    """Row of global hypothesis `row` inside the padded [world, capacity] gather buffer."""
    base, rem = divmod(n, world)
    if row < rem * (base + 1):
        rank = row // (base + 1)
        start = rank * (base + 1)
    else:
        rank = rem + (row - rem * (base + 1)) // max(base, 1)
        start = rem * (base + 1) + (rank - rem) * base
    return rank * shard_capacity(n, world) + (row - start)


def gather_handles(handle: bytes, dist=None, device=None):
    """All ranks' exchange handles in rank order (the list fpb200_exchange_connect wants).  `dist` is an
    initialised torch.distributed (gloo or nccl); the 64 bytes travel as a uint8 tensor, on `device` for nccl."""
    import torch
    if dist is None:
        import torch.distributed as dist
    world = dist.get_world_size()
    mine = torch.tensor(list(handle), dtype=torch.uint8, device=device)
    out = torch.empty(world * len(handle), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, mine)
    flat = bytes(out.cpu().tolist())
    return [flat[i * len(handle):(i + 1) * len(handle)] for i in range(world)]
