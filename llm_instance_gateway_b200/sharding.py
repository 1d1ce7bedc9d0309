"""Multi-GPU plumbing: the request batch shards by request, the snapshot is replicated.

Decisions are independent given a frozen snapshot (Scheduler.Schedule never mutates pod
metrics, pkg/ext-proc/scheduling/scheduler.go:113-122), so the only exchange step is one
broadcast of the packed snapshot blob per refresh tick; picks need no collective — each rank
owns a contiguous slice of the result.  torch.distributed is the plumbing (NCCL on GPUs, gloo in
the CPU tests); no torch types cross the C ABI (raw pointers only).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous request shard [lo, hi) of rank `rank`: result order == request order."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    return total * rank // world, total * (rank + 1) // world


def broadcast_snapshot(blob: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Replicate the packed snapshot (uint8 tensor, same size on every rank) from `src`.

    On GPUs this is one ncclBroadcast over NVLink/NVSwitch of <= ~0.6 MB (P=4096, A=1024)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(blob, src=src)
    return blob


def max_over_ranks(value: float, device) -> float:
    """Multi-GPU timings are reported as the max over ranks."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
