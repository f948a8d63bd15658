"""Host-side multi-GPU plumbing: one process per GPU, torch.distributed for the control plane.

The KNN path shards without a data-path collective: frames are dealt to ranks, queries of
one cloud are split into contiguous ranges.  The one real exchange is the ICP normal-equation
all-reduce, which the C library issues itself through NCCL once the ranks share a communicator id;
the id travels over whatever control plane the host has (here: torch.distributed).

Everything in this file is backend-agnostic (works on gloo/CPU), so it is unit-tested without a GPU.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple


def deal_frames(n_frames: int, rank: int, world: int) -> List[int]:
    """Round-robin frame ownership (BASELINE config 5: 100 frames over 8 GPUs -> 12 or 13 each)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_frames, world))


def split_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [begin, end) slice of n queries for `rank` (sizes differ by at most 1)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def _tensor(value: float, device):
    import torch

    return torch.tensor([float(value)], dtype=torch.float64, device=device)


def max_over_ranks(dist, value: float, device="cpu") -> float:
    """Timing rule of the bench contract: a multi-GPU duration is the MAX over ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = _tensor(value, device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, value: float, device="cpu") -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = _tensor(value, device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def share_unique_id(dist, make_id: Callable[[], bytes], device="cpu") -> bytes:
    """Rank 0 creates the NCCL unique id (pcr_comm_unique_id), every rank receives the same bytes."""
    import torch

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return make_id()
    n = 128
    if dist.get_rank() == 0:
        raw = make_id()
        if len(raw) != n:
            raise ValueError("unique id must be 128 bytes")
        t = torch.tensor(list(raw), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(n, dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def gather_frame_results(dist, local: Sequence, n_frames: int, rank: int, world: int) -> Optional[list]:
    """Reassemble per-frame results dealt with deal_frames() on rank 0 (control plane only)."""
    if dist is None or not dist.is_initialized() or world == 1:
        return list(local)
    out = [None] * world
    dist.all_gather_object(out, list(local))
    if rank != 0:
        return None
    merged = [None] * n_frames
    for r in range(world):
        for j, f in enumerate(deal_frames(n_frames, r, world)):
            merged[f] = out[r][j]
    return merged
