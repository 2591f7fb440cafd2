"""Multi-GPU = independent replicas: one process per GPU, each with its own engine and its own
requests; nothing is exchanged on the data path.  The only collective is the measurement itself: a barrier, the MAX
of the per-rank wall time and the SUM of the per-rank work."""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def shard_requests(n_requests: int, rank: int, world_size: int) -> List[int]:
    """Round-robin request partition: request i is served by rank i % world_size."""
    return [i for i in range(n_requests) if i % world_size == rank]


def aggregate(times_ms: Sequence[float], work: Sequence[float], device=None) -> Tuple[List[float], List[float]]:
    """MAX over ranks of each time, SUM over ranks of each work item (identity when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(times_ms), list(work)
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    w = torch.tensor(list(work), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(w, op=dist.ReduceOp.SUM)
    return t.tolist(), w.tolist()


def rtf(frames: float, ms: float, frame_seconds: float = 0.08) -> float:
    return frames * frame_seconds / (ms / 1000.0) if ms > 0 else 0.0
