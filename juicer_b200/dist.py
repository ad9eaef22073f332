"""Multi-GPU plumbing: one process per GPU, utterances sharded across ranks.

Decodes are independent (they share only the read-only network and models), so the path
shards by utterance with NO data-path collective: every rank builds
the same device tables and decodes what it claims from one utterance list shared through a node-local
atomic counter (whole-utterance work stealing, jgpu_queue_* / jgpu_decode_queue of the C ABI);
torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for barriers, to gather the
fixed-size result records and to reduce the timing (max over ranks) and frame counts (sum).
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


_QUEUE_GENERATION = 0


class SharedQueue:
    """The node-local utterance counter of the C ABI (jgpu_queue_*, juicer_b200/csrc/host_queue.cpp): one 64-bit
    atomic in POSIX shared memory.  `SharedQueue.collective()` is called by all ranks of the process group together:
    rank 0 creates a FRESH segment (its name carries a generation number that advances with every call, so two
    queues never share a counter), the others open it after a barrier.  Without a process group the queue is
    private to the process."""

    def __init__(self, name: str, create: bool):
        from . import api
        import ctypes as C
        self.lib = api.load_library()
        self.lib.jgpu_queue_open.argtypes = [C.c_char_p, C.c_int32, C.POINTER(C.c_void_p)]
        self.lib.jgpu_queue_reset.argtypes = [C.c_void_p]
        self.lib.jgpu_queue_claim.argtypes = [C.c_void_p, C.c_int64]
        self.lib.jgpu_queue_claim.restype = C.c_int64
        self.lib.jgpu_queue_position.argtypes = [C.c_void_p]
        self.lib.jgpu_queue_position.restype = C.c_int64
        self.lib.jgpu_queue_close.argtypes = [C.c_void_p]
        self.q = C.c_void_p()
        rc = self.lib.jgpu_queue_open(name.encode(), int(create), C.byref(self.q))
        if rc < 0:
            raise api.JuicerError(f"jgpu_queue_open({name}) failed ({rc}): {self.lib.jgpu_last_error().decode(errors='replace')}")
        self.name = name

    @classmethod
    def collective(cls, tag: str = "utts") -> "SharedQueue":
        global _QUEUE_GENERATION
        _QUEUE_GENERATION += 1
        rank, world = 0, 1
        dist = None
        try:
            import torch.distributed as dist_
            if dist_.is_initialized():
                dist, rank, world = dist_, dist_.get_rank(), dist_.get_world_size()
        except ImportError:
            pass
        port = os.environ.get("MASTER_PORT", "0") if world > 1 else f"p{os.getpid()}"
        name = f"juicer_b200.{port}.{tag}.{_QUEUE_GENERATION}"
        if world == 1:
            return cls(name, True)
        q = cls(name, True) if rank == 0 else None
        dist.barrier()
        if q is None:
            q = cls(name, False)
        dist.barrier()
        return q

    def claim(self, n: int = 1) -> int:
        return int(self.lib.jgpu_queue_claim(self.q, n))

    def reset(self) -> None:
        self.lib.jgpu_queue_reset(self.q)

    @property
    def position(self) -> int:
        return int(self.lib.jgpu_queue_position(self.q))

    def close(self) -> None:
        if getattr(self, "q", None):
            self.lib.jgpu_queue_close(self.q)
            self.q = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class UtteranceQueue:
    """Whole-utterance work stealing across ranks at the granularity of `wave` utterances:
    the list, sorted longest first, is one shared queue, and a rank claims its next wave with one atomic fetch-add
    on a SharedQueue — a rank that finishes early simply keeps claiming what a static split would have left to a
    slower one.  (The CUDA decoder claims utterance by utterance as its lanes run dry: api.WFSTDecoderLite.
    decode_queue.)  Only the counter crosses ranks; results are gathered at the end (`gather_results`).  Every
    instance gets a fresh counter."""

    def __init__(self, n_frames: Sequence[int], wave: int, name: str = "utts"):
        self.order = sorted(range(len(n_frames)), key=lambda i: (-int(n_frames[i]), i))
        self.wave = max(int(wave), 1)
        self.shared = SharedQueue.collective(name)

    def claim(self) -> List[int]:
        """Next wave of utterance indices for the calling rank; empty when the queue is drained."""
        start = self.shared.claim(self.wave)
        return self.order[start:start + self.wave] if start < len(self.order) else []

    def close(self) -> None:
        self.shared.close()


def decode_with_stealing(decode_wave, n_frames: Sequence[int], wave: int, name: str = "utts") -> Dict[int, dict]:
    """Drains the shared queue: `decode_wave(indices) -> {index: record}` is called with one wave at a
    time.  Returns this rank's records."""
    q = UtteranceQueue(n_frames, wave, name)
    local: Dict[int, dict] = {}
    while True:
        idx = q.claim()
        if not idx:
            break
        local.update(decode_wave(idx))
    try:
        import torch.distributed as dist
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.barrier()                                   # nobody unlinks the segment while another rank still claims
    except ImportError:
        pass
    q.close()
    return local


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
