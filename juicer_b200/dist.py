"""Multi-GPU plumbing: one process per GPU, utterances sharded across ranks.

Decodes are independent (they share only the read-only network and models), so the path
shards by utterance with NO data-path collective: every rank builds
the same device tables, decodes its own shard, and torch.distributed (NCCL on GPUs, gloo in
the CPU tests) is used only to gather the fixed-size result records and to reduce the
timing (max over ranks) and frame counts (sum).
"""
from __future__ import annotations

import os
from typing import Dict, List, Sequence, Tuple


def shard_utterances(n_frames: Sequence[int], world: int) -> List[List[int]]:
    """Longest-processing-time assignment: utterance indices per rank, balanced by frame count.
    Deterministic, identical on every rank."""
    order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
    load = [0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        shards[r].append(i)
        load[r] += int(n_frames[i]) + 1
    return shards


def env_rank_world() -> Tuple[int, int, int]:
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init_process_group(backend: str):
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend)
    return dist


def gather_results(local: Dict[int, dict], n_total: int) -> List[dict]:
    """All ranks contribute {utterance index: result record}; every rank gets the full list."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        merged = dict(local)
    else:
        parts: List[Dict[int, dict]] = [None] * dist.get_world_size()   # type: ignore[list-item]
        dist.all_gather_object(parts, local)
        merged = {}
        for p in parts:
            merged.update(p)
    missing = [i for i in range(n_total) if i not in merged]
    if missing:
        raise RuntimeError(f"utterances without a result: {missing[:8]}")
    return [merged[i] for i in range(n_total)]


def reduce_time_and_frames(ms: float, frames: int, device=None) -> Tuple[float, int]:
    """(max over ranks of elapsed ms, sum over ranks of frames)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return ms, frames
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    f = torch.tensor([frames], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(f, op=dist.ReduceOp.SUM)
    return float(t.item()), int(f.item())
