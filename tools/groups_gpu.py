"""Experiment: several decoder handles (lane groups) on one GPU, each on its own stream, decoding disjoint halves of the
batch concurrently from host threads — do the latency-bound small kernels and launch gaps of one group hide behind the
other group's big kernels?      python tools/groups_gpu.py [--groups 1,2,3] [--lanes 256]"""
import argparse, json, os, sys, threading, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from juicer_b200 import api

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--groups", default="1,2")
    ap.add_argument("--lanes", type=int, default=256, help="lanes in all (split over the groups)")
    ap.add_argument("--utts", type=int, default=512)
    ap.add_argument("--workload", default="c3")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    m, net, tee, kw, files = bench.build_fixture(args.workload, "/tmp/juicer_b200_bench", 0)
    network = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
    models = api.HTKFlatModels(files["jmbi"])
    feats = bench.sample_utterances(net, m, tee, args.utts, 300, 1000, seed=1000)
    order = np.argsort([-f.shape[0] for f in feats], kind="stable")
    out = []
    for G in [int(x) for x in args.groups.split(",")]:
        decs, batches, streams = [], [], []
        for g in range(G):
            d = api.WFSTDecoderLite(network, models, 0.0, kw["main_beam"], 0.0, 0.0, 0, n_lanes=args.lanes // G, device=0)
            s = torch.cuda.Stream(device=dev)
            d.set_stream(s.cuda_stream)
            sub = [feats[int(u)] for u in order[g::G]]           # interleaved by length: equal work per group
            decs.append(d); streams.append(s); batches.append(bench.Batch(sub, dev, torch))
        def run(g):
            decs[g].decode_batch_device(batches[g].packed_dev.data_ptr(), batches[g].offsets, batches[g].n_frames, want_results=False)
        def step():
            th = [threading.Thread(target=run, args=(g,)) for g in range(G)]
            for t in th: t.start()
            for t in th: t.join()
        step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            step()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / args.reps
        rows = sum(b.rows for b in batches)
        rec = {"groups": G, "lanes_per_group": args.lanes // G, "frames_per_s": rows / dt, "s_per_step": dt}
        print(json.dumps(rec), flush=True)
        out.append(rec)
        for d in decs: d.close()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/groups.json", "w"), indent=1)

if __name__ == "__main__":
    main()
