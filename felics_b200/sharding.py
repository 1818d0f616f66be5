"""Multi-GPU sharding of image batches.

Images are independent units: contiguous ranges of the batch go to the ranks, every rank
encodes its range into its own arena, and the only exchange is a gather of the per-image
compressed sizes so that any rank can place image i in the global stream.  There is no
collective on the data path (no NCCL on the hot path, as BASELINE.json's north_star states).
A single image cannot be split bit-exactly (one estimator and one bit chain per plane).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """(first, count) of the contiguous slice of `n` images owned by `rank`; the first
    n % world ranks hold one extra image."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def gather_sizes(local_sizes: Sequence[int], group=None) -> np.ndarray:
    """All ranks' per-image sizes, concatenated in rank order (uint64).  Uses the current
    torch.distributed process group (gloo on CPU, nccl on GPU); a single process returns
    its own sizes."""
    import torch
    import torch.distributed as dist

    local = np.asarray(local_sizes, dtype=np.int64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return local.astype(np.uint64)
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(local)], dtype=torch.int64, device=dev), group=group)
    counts = [int(c.item()) for c in counts]
    width = max(counts + [1])
    mine = torch.zeros(width, dtype=torch.int64, device=dev)
    mine[: len(local)] = torch.from_numpy(local).to(dev)
    parts = [torch.zeros(width, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    return np.concatenate([p[:c].cpu().numpy() for p, c in zip(parts, counts)]).astype(np.uint64)


def global_offsets(all_sizes: np.ndarray) -> np.ndarray:
    """offsets[n+1] of the concatenated .fel stream from all ranks' per-image sizes."""
    out = np.zeros(len(all_sizes) + 1, dtype=np.uint64)
    np.cumsum(np.asarray(all_sizes, dtype=np.uint64), out=out[1:])
    return out


def reduce_scalar(value: float, op: str = "max", group=None, device: Optional[object] = None) -> float:
    """max / sum of a python float over the ranks (timing: max over ranks; work: sum)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM, group=group)
    return float(t.item())
