"""CPU suite: the N>1 path (utterance sharding + result gather + timing reduction) with
world_size 2 over gloo.  The per-rank decoder here is the oracle port — the sharding logic is
what is under test; the CUDA decoder takes its place on the GPU box."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import Golden, flat_tables_from_files

from juicer_b200 import _abi, dist as jdist


def test_shard_balance_and_coverage():
    rng = np.random.default_rng(0)
    n = rng.integers(1, 1000, size=257)
    for world in (1, 2, 3, 8):
        sh = jdist.shard_utterances(n, world)
        flat = sorted(i for s in sh for i in s)
        assert flat == list(range(len(n)))
        loads = [sum(int(n[i]) + 1 for i in s) for s in sh]
        assert max(loads) - min(loads) <= int(n.max()) + 1
    assert jdist.shard_utterances([], 4) == [[], [], [], []]
    assert jdist.shard_utterances([5], 2) == [[0], []]


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from oracle.binding import OraclePort
    d = jdist.init_process_group("gloo")
    g = Golden("mixed")
    tabs, _n, _m = flat_tables_from_files(g.files)
    p = OraclePort(tabs, _abi.make_cfg(**g.kw))
    feats = [g.feats(u) for u in range(g.n_utts)] + [g.feats(0)[:37]]
    shards = jdist.shard_utterances([f.shape[0] for f in feats], world)
    local = {}
    for u in shards[rank]:
        r = p.decode(feats[u])
        local[u] = dict(status=r.status, labels=r.labels, times=r.times, totals=r.totals.view(np.uint32).tolist())
    allr = jdist.gather_results(local, len(feats))
    ms, frames = jdist.reduce_time_and_frames(10.0 * (rank + 1), sum(feats[u].shape[0] for u in shards[rank]))
    if rank == 0:
        q.put((allr, ms, frames))
    d.barrier()
    d.destroy_process_group()


def test_two_rank_gloo_matches_single_process(oracle_port_lib, product_lib):
    from oracle.binding import OraclePort
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allr, ms, frames = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = Golden("mixed")
    tabs, _n, _m = flat_tables_from_files(g.files)
    single = OraclePort(tabs, _abi.make_cfg(**g.kw))
    feats = [g.feats(u) for u in range(g.n_utts)] + [g.feats(0)[:37]]
    assert frames == sum(f.shape[0] for f in feats) and ms == 20.0
    for u, f in enumerate(feats):
        r = single.decode(f)
        assert allr[u]["status"] == r.status and allr[u]["labels"] == r.labels and allr[u]["times"] == r.times
        assert allr[u]["totals"] == r.totals.view(np.uint32).tolist()


def _steal_worker(rank, world, port, q):
    import time
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    d = jdist.init_process_group("gloo")
    n_frames = list(range(40, 0, -1))                       # 40 utterances, longest first already
    claimed = []

    def decode_wave(idx):
        claimed.append(list(idx))
        time.sleep(0.05 if rank == 1 else 0.0)              # rank 1 is the slow GPU
        return {u: dict(rank=rank, frames=n_frames[u]) for u in idx}

    local = jdist.decode_with_stealing(decode_wave, n_frames, wave=3)
    allr = jdist.gather_results(local, len(n_frames))
    # a second list through a second queue with the same (default) name: a fresh counter, nothing is skipped
    # (the store key of the first version was never reset)
    local2 = jdist.decode_with_stealing(lambda idx: {u: dict(rank=rank, frames=n_frames[u]) for u in idx}, n_frames, wave=5)
    allr2 = jdist.gather_results(local2, len(n_frames))
    assert [r["frames"] for r in allr2] == n_frames
    if rank == 0:
        q.put(allr)
    d.barrier()
    d.destroy_process_group()


def test_work_stealing_queue_two_ranks(product_lib):
    """Every utterance is decoded exactly once, and the fast rank ends up with more of them."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_steal_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    allr = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r["frames"] for r in allr] == list(range(40, 0, -1))
    by_rank = [sum(1 for r in allr if r["rank"] == k) for k in range(2)]
    assert sum(by_rank) == 40 and by_rank[0] > by_rank[1]


def test_queue_without_process_group_is_a_plain_loop(product_lib):
    q = jdist.UtteranceQueue([5, 9, 1, 7], wave=3)
    assert q.claim() == [1, 3, 0] and q.claim() == [2] and q.claim() == []
    q.close()
    q2 = jdist.UtteranceQueue([5, 9, 1, 7], wave=3)          # same default name, fresh counter
    assert q2.claim() == [1, 3, 0]
    q2.close()


def _claim_worker(name, n, out_q):
    q = jdist.SharedQueue(name, False)
    mine = []
    while True:
        i = q.claim(1)
        if i >= n:
            break
        mine.append(i)
    out_q.put(mine)
    q.close()


def test_shared_queue_hands_out_every_position_once(product_lib):
    """The C-ABI counter (jgpu_queue_*: POSIX shared memory + atomic fetch-add) under three concurrent processes."""
    name = f"juicer_b200.test.{os.getpid()}"
    owner = jdist.SharedQueue(name, True)
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    n = 20000
    procs = [ctx.Process(target=_claim_worker, args=(name, n, out_q)) for _ in range(3)]
    for p in procs:
        p.start()
    got = [out_q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    flat = sorted(i for g in got for i in g)
    assert flat == list(range(n))
    assert owner.position >= n
    owner.reset()
    assert owner.claim(4) == 0 and owner.claim(1) == 4
    owner.close()
    with pytest.raises(Exception):
        jdist.SharedQueue(name, False)                       # the creator unlinked the segment
