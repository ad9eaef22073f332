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


class UtteranceQueue:
    """Whole-utterance work stealing across ranks: the utterance list, sorted
    longest first, is one shared queue; a rank claims the next `wave` utterances with an atomic
    fetch-add on the process group's key-value store, so a rank that finishes early simply keeps
    claiming what a static split would have left to a slower one.  Only the counter crosses ranks:
    every rank can read every utterance's features (host side), and results are gathered at the end
    (`gather_results`), so there is still no data-path collective.  Without a process group the
    counter is local and the queue degenerates to a plain loop."""

    def __init__(self, n_frames: Sequence[int], wave: int, name: str = "juicer_b200/utt_queue"):
        self.order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
        self.wave = max(int(wave), 1)
        self.key = name
        self._local = 0
        self._store = None
        try:
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_world_size() > 1:
                from torch.distributed.distributed_c10d import _get_default_store
                self._store = _get_default_store()
        except ImportError:
            pass

    def claim(self) -> List[int]:
        """Next wave of utterance indices for the calling rank; empty when the queue is drained."""
        if self._store is not None:
            start = int(self._store.add(self.key, self.wave)) - self.wave
        else:
            start = self._local
            self._local += self.wave
        return self.order[start:start + self.wave] if start < len(self.order) else []


def decode_with_stealing(decode_wave, n_frames: Sequence[int], wave: int, name: str = "juicer_b200/utt_queue") -> Dict[int, dict]:
    """Drains the shared queue: `decode_wave(indices) -> {index: record}` is called with one wave at a
    time (on the GPU: one lock-step batch of the rank's decoder).  Returns this rank's records."""
    q = UtteranceQueue(n_frames, wave, name)
    local: Dict[int, dict] = {}
    while True:
        idx = q.claim()
        if not idx:
            return local
        local.update(decode_wave(idx))


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
