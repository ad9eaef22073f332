"""BASELINE configs[4]: 64k-vocab trigram-shaped network (~1.45M states / ~5.9M arcs), main beam swept over
150..400 with and without histogram pruning, one GPU.

    python tools/c5_sweep.py [--lanes 128] [--utts 256] [--out gpurun_out/c5_sweep.json]

The network / models / utterances are built once; every (beam, max_hyps) setting gets its own decoder (the
pruning settings are constructor arguments, as in WFSTDecoderLite) and decodes the same batch resident in
HBM, timed with CUDA events on the decoder's stream after one warm-up pass.  Reported per setting: frames/s,
the reference's work counters per frame (active models / emitting hypotheses / end hypotheses) and how many
utterances still recover the planted word sequence."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juicer_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lanes", type=int, default=128)
    ap.add_argument("--utts", type=int, default=256)
    ap.add_argument("--min-frames", type=int, default=300)
    ap.add_argument("--max-frames", type=int, default=1000)
    ap.add_argument("--beams", default="150,200,250,300,350,400")
    ap.add_argument("--max-hyps", default="0,6000,20000")
    ap.add_argument("--workdir", default="/tmp/juicer_b200_bench/c5_sweep")
    ap.add_argument("--out", default="gpurun_out/c5_sweep.json")
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    t0 = time.time()
    m, net, tee, kw = synth.named_config("c5")
    files = synth.make_fixture("c5", args.workdir, m, net)
    network = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
    models = api.HTKFlatModels(files["jmbi"])
    ps = synth.PathSampler(net, m, tee_hmms=tee)
    rng = np.random.default_rng(5005)
    samples = [ps.sample(int(rng.integers(args.min_frames, args.max_frames + 1)), rng) for _ in range(args.utts)]
    feats = [s[0] for s in samples]
    truth = [list(s[1]) for s in samples]
    n_frames = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
    offsets = np.concatenate([[0], np.cumsum(n_frames)[:-1]]).astype(np.int64)
    packed = torch.from_numpy(np.concatenate(feats, axis=0)).to(dev)
    rows = int(n_frames.sum())
    print(f"c5: {network.c.n_states} states / {network.c.n_arcs} arcs, {args.utts} utterances, {rows} frames "
          f"(set-up {time.time() - t0:.0f} s)", flush=True)
    stream = torch.cuda.Stream(device=dev)
    out = []
    for beam in [float(b) for b in args.beams.split(",")]:
        for mh in [int(x) for x in args.max_hyps.split(",")]:
            dec = api.WFSTDecoderLite(network, models, 0.0, beam, 0.0, 0.0, mh, n_lanes=args.lanes, device=0)
            dec.set_stream(stream.cuda_stream)
            dec.decode_batch_device(packed.data_ptr(), offsets, n_frames, want_results=False)      # warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            res = dec.decode_batch_device(packed.data_ptr(), offsets, n_frames, want_results=True)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            st = dec.stats(-1)
            nf = max(st["n_frames"], 1)
            ok = sum(1 for r, t in zip(res, truth) if r.status > 0 and r.labels == t)
            failed = sum(1 for r in res if r.status <= -10)
            # status = JGPU_E_CAPACITY - 10 - (error bits << 8): 1 active list, 2 arrivals, 4 word-boundary arena, 16 hub list
            err_bits = sorted({(-(r.status + 13)) >> 8 for r in res if r.status <= -10})
            rec = {"main_beam": beam, "max_hyps": mh, "frames_per_s": rows / (ms * 1e-3), "ms": ms,
                   "active_models_per_frame": st["total_active_models"] / nf,
                   "emit_hyps_per_frame": st["total_active_emit_hyps"] / nf,
                   "end_hyps_per_frame": st["total_active_end_hyps"] / nf,
                   "planted_sequence_recovered": ok, "capacity_failures": failed, "second_passes": dec.retry_count, "capacity_error_bits": err_bits, "utterances": args.utts,
                   "lanes": args.lanes}
            out.append(rec)
            print(json.dumps(rec), flush=True)
            dec.close()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    json.dump({"workload": "c5", "states": int(network.c.n_states), "arcs": int(network.c.n_arcs), "frames": rows,
               "results": out}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
