"""One-process-per-GPU plumbing (torch.distributed): rank -> column band, max/sum reductions of
timings and ray counters, and a host-side stitch of the bands for verification.

The render path itself has NO collective: pixels are independent, the scene is
replicated, each rank's finished band goes straight to host memory.  torch.distributed is used
only for the barrier around the timed region and to combine per-rank numbers.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np

from .partition import column_band


def env_world() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: Optional[str] = None):
    """Initialise the default process group when launched under torchrun; returns (rank, world, local_rank)."""
    rank, world, local = env_world()
    if world > 1:
        import torch.distributed as dist

        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", "29511")
            dist.init_process_group(backend=backend or "gloo", rank=rank, world_size=world)
    return rank, world, local


def shutdown() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()


def barrier() -> None:
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def _reduce(value: float, op_name: str, device=None) -> float:
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op_name))
    return float(t.item())


def gather_floats(value: float, device=None) -> list[float]:
    """Every rank's value, in rank order (diagnostics: per-rank kernel times)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return [float(value)]
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def broadcast_object(obj, src: int = 0):
    """Setup-time only (e.g. the band cut computed on rank 0); never on the render path."""
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def reduce_max(value: float, device=None) -> float:
    return _reduce(value, "MAX", device)


def reduce_sum(value: float, device=None) -> float:
    return _reduce(value, "SUM", device)


def rank_band(width: int) -> tuple[int, int]:
    rank, world, _ = env_world()
    return column_band(width, rank, world)


def stitch_bands(band: np.ndarray, width: int) -> Optional[np.ndarray]:
    """Verification helper: concatenates every rank's band [x0:x1, H, 3] on rank 0 (host side,
    through the process group's object gather).  Not on the measured path."""
    import torch.distributed as dist

    rank, world, _ = env_world()
    if world == 1:
        return band
    parts = [None] * world if rank == 0 else None
    dist.gather_object(np.ascontiguousarray(band), parts, dst=0)
    if rank != 0:
        return None
    full = np.concatenate(parts, axis=0)
    assert full.shape[0] == width
    return full
