"""GPU suite (-m gpu): BASELINE configs[3]'s mechanism — several ranks decode ONE utterance list, each taking the
next utterance from the shared counter when a lane runs dry (jgpu_decode_queue).  Two processes (two GPUs when the
box has them, else both on GPU 0), gloo for the barriers and the gather; results must equal the golden vectors of
the reference and every utterance must be decoded exactly once."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from helpers import Golden, bits, flat_tables_from_files

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch
    from juicer_b200 import api, dist as jdist
    d = jdist.init_process_group("gloo")
    dev = rank % torch.cuda.device_count()
    g = Golden("mixed")
    tabs, net, models = flat_tables_from_files(g.files)
    kw = g.kw
    dec = api.WFSTDecoderLite(net, models, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                              kw.get("word_beam", 0.0), kw.get("max_hyps", 0), n_lanes=3, device=dev)
    feats = [g.feats(u % 3) for u in range(40)]
    dec.decode_batch(feats[:3])                            # module load and graph capture are not part of the race
    records = {}
    for mode in ("host", "device"):
        q = jdist.SharedQueue.collective("gputest")
        if mode == "host":
            mine, busy = dec.decode_queue(q, feats)
        else:
            nfr = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
            off = np.concatenate([[0], np.cumsum(nfr)[:-1]]).astype(np.int64)
            packed = torch.from_numpy(np.concatenate(feats, axis=0)).to(f"cuda:{dev}")
            mine, busy = dec.decode_queue_device(q, packed.data_ptr(), off, nfr)
        d.barrier()
        q.close()
        local = {u: dict(rank=rank, status=r.status, labels=r.labels, times=r.times,
                         totals=np.asarray(r.totals, dtype=np.float32).view(np.uint32).tolist()) for u, r in mine.items()}
        records[mode] = (jdist.gather_results(local, len(feats)), busy)
    if rank == 0:
        out_q.put(records)
    d.barrier()
    dec.close()
    d.destroy_process_group()


def test_two_ranks_share_one_utterance_queue(oracle_port_lib, product_lib):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out_q)) for r in range(2)]
    for p in procs:
        p.start()
    records = out_q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = Golden("mixed")
    for mode in ("host", "device"):
        allr, busy = records[mode]
        assert len(allr) == 40 and busy > 0.0
        for u, r in enumerate(allr):
            z = g.z
            k = u % 3
            assert r["status"] == int(z[f"status{k}"]) and r["labels"] == z[f"labels{k}"].tolist(), (mode, u, r)
            assert r["times"] == z[f"times{k}"].tolist()
            assert np.array_equal(np.asarray(r["totals"], dtype=np.uint32), z[f"totals{k}"]), (mode, u)
        assert {r["rank"] for r in allr} == {0, 1}, "both ranks must have claimed utterances"
