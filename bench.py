#!/usr/bin/env python
"""bench.py — decoded frames/sec of the WFST token-passing decode path on B200.

    python bench.py --gpus N --steps K --warmup W                 (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W  (reference CPU path)

A "step" decodes one batch of synthetic utterances (39-d MFCC-like frames sampled along
random accepted paths) on the BASELINE.json configs[2]-shaped network ("c3": 20k-word
trigram-shaped C.L.G, ~440k states / ~1.8M arcs, 6000 tied 16-mix GMMs, main beam 250).
One process per GPU; utterances are independent, so ranks decode disjoint batches of the
same size (weak scaling) with no data-path collective; torch.distributed only reduces the
timing (max over ranks) and the frame counts (sum).

Prints ONE JSON line (rank 0).  `value` = frames/s with features resident in HBM, timed with
CUDA events on the decoder's stream; `e2e` = the same batches through jgpu_decode_batch with
pinned HOST buffers (H2D feature copies and D2H result copies inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from juicer_b200 import synth  # noqa: E402

WORKLOADS = {
    # name: (named_config, description)
    "c3": "20k-word trigram-shaped C.L.G (~440k states / ~1.8M arcs), 4000 HMMs over 6000 tied 16-mix GMMs, "
          "main beam 250, utterances of 300-1000 frames",
    "c2": "1k-word bigram C.L.G (~42k states / ~51k arcs), 2000 triphone HMMs x 16-mix, main beam 200, "
          "utterances of 300-1000 frames",
    "c3s": "c3 topology at 1/8 scale (smoke runs)",
    "c5": "64k-word trigram-shaped C.L.G (~1.45M states / ~5.9M arcs), 4000 HMMs over 6000 tied 16-mix GMMs, "
          "utterances of 300-1000 frames; main beam / histogram limit from --beam / --max-hyps",
}
METRIC = "decoded frames/sec (xRT) on composed H∘C∘L∘G WFST"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="graft", choices=["graft", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--utts", type=int, default=512, help="utterances per step per GPU")
    ap.add_argument("--lanes", type=int, default=256, help="utterances decoded in lock-step")
    ap.add_argument("--min-frames", type=int, default=300)
    ap.add_argument("--max-frames", type=int, default=1000)
    ap.add_argument("--cpu-sample-utts", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--beam", type=float, default=0.0, help="override the workload's main beam (beam sweeps, BASELINE configs[4])")
    ap.add_argument("--max-hyps", type=int, default=-1, help="override the histogram-pruning limit (0 = off)")
    ap.add_argument("--workdir", default=os.environ.get("JUICER_BENCH_DIR", "/tmp/juicer_b200_bench"))
    return ap.parse_args()


_OVERRIDES: Dict[str, float] = {}


def build_fixture(workload: str, workdir: str, rank: int):
    m, net, tee, kw = synth.named_config(workload)
    kw = dict(kw, **_OVERRIDES)
    d = os.path.join(workdir, f"{workload}_r{rank}")
    files = synth.make_fixture(workload, d, m, net)
    return m, net, tee, kw, files


def sample_utterances(net, m, tee, n, lo, hi, seed) -> List[np.ndarray]:
    ps = synth.PathSampler(net, m, tee_hmms=tee)
    rng = np.random.default_rng(seed)
    return [ps.sample(int(rng.integers(lo, hi + 1)), rng)[0] for _ in range(n)]


# ---------------------------------------------------------------------------------------
# clocks (profiling recipe: sample DURING the timed region)
# ---------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in out.strip().split(",")]
                if len(p) >= 6:
                    self.rows.append(p)
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self) -> Dict:
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------
# reference arm: the reference CPU decoder on the host cores
# ---------------------------------------------------------------------------------------
_REF_DEC = None


def _ref_init(files, kw, use_ref):
    """Pool initializer: every worker process loads the network and models once."""
    global _REF_DEC
    sys.path.insert(0, ROOT)
    from oracle.binding import OraclePort, OracleRef
    if use_ref:
        _REF_DEC = OracleRef(files, **kw)
    else:
        from juicer_b200 import _abi, api
        net = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
        models = api.HTKFlatModels(files["jmbi"])
        _REF_DEC = OraclePort(_abi.FlatTables(net.arrays(), net.init_state, models.arrays()), _abi.make_cfg(**kw))
        _REF_DEC._keep = (net, models)


def _ref_worker(x):
    t0 = time.perf_counter()
    _REF_DEC.decode(x)
    return x.shape[0], time.perf_counter() - t0


def run_reference(args) -> None:
    """CPU baseline = the reference's own WFSTDecoderLite/HTKFlatModels objects (oracle/_ref, kind
    "reference") when the prebuilt library is present, else the plain-C restatement (kind "port").
    The reference is single-threaded; "all the host threads it can use" = one independent process
    per core over a disjoint split of the utterances."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    from oracle import binding
    binding.build(ref=True, port=True)
    use_ref = os.path.exists(binding.REF_SO)
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    m, net, tee, kw, files = build_fixture(args.workload, args.workdir, 0)
    # bounded sample: one short utterance per core per step (~3-4 s of CPU work each on c3)
    lo, hi = args.min_frames, args.min_frames + 20
    feats = sample_utterances(net, m, tee, cores, lo, hi, seed=12345)
    ctx = mp.get_context("spawn")
    times = []
    with ctx.Pool(cores, initializer=_ref_init, initargs=(files, kw, use_ref)) as pool:
        pool.map(_ref_worker, [f[:2] for f in feats], chunksize=1)        # force every worker to finish loading
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, feats, chunksize=1)
            dt = time.perf_counter() - t0                                  # wall time of the step (all cores busy)
            if step >= args.warmup:
                times.append((sum(r[0] for r in res), dt, max(r[1] for r in res)))
    frames = sum(t[0] for t in times)
    sec = sum(t[1] for t in times)
    value = frames / sec
    kind = "reference" if use_ref else "port"
    sample = (f"{cores} utterances of {lo}-{hi} frames per step, one per host core "
              f"({'unmodified reference objects, g++ -O2' if use_ref else 'plain-C restatement, gcc -O2'})")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / max(len(times), 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload]}", "decoder": kw, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "xrt": value / 100.0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------
# roofline bookkeeping
# ---------------------------------------------------------------------------------------
def algorithmic_bytes(stats: Dict[str, int], dims: Dict[str, int], n_rows: int) -> Dict[str, float]:
    """Algorithmic bytes moved by each kernel over one whole step, from the decoder's own work
    counters and the device record sizes of DESIGN.md ("Kernels"): instance record 16 B, token 16 B,
    arrival record 32 B, state row 16 B, arc 16 B, slotmap entry 4 B, word-boundary record 32 B.
    Minimal traffic: every record counted once."""
    A = stats["total_active_models"]          # instances walked by the internal phase (sum over frames)
    H = stats["total_active_emit_hyps"]       # live emitting tokens after the internal phase
    Ea = stats["total_active_end_hyps"]       # live exit tokens = arrival records written by k_internal
    E = stats["total_proc_end_hyps"]          # exit tokens passing the end/word beam = first-round arrivals
    X = stats["total_arcs_expanded"]          # out-arcs of the states committed (hub rows included)
    W = stats["total_entry_writes"]           # distinct destination arcs whose entry token was written
    P = stats["total_paths"]                  # word-boundary records appended
    D, G, M = dims["D"], dims["n_gmm"], dims["C"]
    return {
        # instance record + entry token read per instance; emitting tokens read + written, one acoustic score
        # per live token; one arrival record per live exit token
        "k_internal": A * (16 + 16) + H * (16 + 16 + 4) + Ea * 32,
        # (only with an end / word beam) arrival records re-read and filtered
        "k_filter": Ea * 16,
        # expansion round 0: arrival record + state row per record, word-boundary records
        "k_expand": E * (32 + 16) + P * 32,
        "k_expand_r1": 0.0, "k_expand_r2": 0.0,   # later rounds: a few hundred records per step
        # commit: arrival record + state row per record, arc + slotmap entry per arc walked, entry token +
        # instance record + slotmap entry per entry written (rows of hub-like states are walked by
        # k_commit_huge; their bytes are counted here)
        "k_commit": E * (32 + 16) + X * (16 + 4) + W * (16 + 16 + 4),
        "k_commit_huge": 0.0,
        "k_boundary": 0.0,
        # one parameter pass per launch (frames x lanes rows share it) + features in, scores out
        "k_gmm_scores": dims["gmm_launches"] * G * M * (2 * D + 1) * 4 + n_rows * (D * 4 + G * 4),
    }


_REAL_STDOUT = None


def claim_stdout() -> None:
    """stdout must carry exactly ONE JSON line, but native libraries write to file descriptor 1 on their own
    (NCCL prints its version banner there whatever NCCL_DEBUG_FILE says): keep a private copy of the real stdout
    for the result line and point fd 1 at stderr for everything else."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: Dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main() -> None:
    args = parse_args()
    claim_stdout()
    if args.beam > 0:
        _OVERRIDES["main_beam"] = args.beam
    if args.max_hyps >= 0:
        _OVERRIDES["max_hyps"] = args.max_hyps
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    from juicer_b200 import api, dist as jdist

    rank, world, local_rank = jdist.env_rank_world()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the decode path has no CPU implementation "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        jdist.init_process_group("nccl")
        import torch.distributed as dist
    api.load_library()

    m, net, tee, kw, files = build_fixture(args.workload, args.workdir, rank)
    network = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
    models = api.HTKFlatModels(files["jmbi"])
    dec = api.WFSTDecoderLite(network, models, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                              kw.get("word_beam", 0.0), kw.get("max_hyps", 0), n_lanes=args.lanes, device=local_rank)
    stream = torch.cuda.Stream(device=dev, priority=-1)      # the search is latency-critical; scoring may run below it
    dec.set_stream(stream.cuda_stream)

    # per-rank batch (weak scaling: every GPU decodes its own `utts` utterances per step)
    feats = sample_utterances(net, m, tee, args.utts, args.min_frames, args.max_frames, seed=1000 + rank)
    n_frames = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
    offsets = np.concatenate([[0], np.cumsum(n_frames)[:-1]]).astype(np.int64)
    rows = int(n_frames.sum())
    packed_host = torch.from_numpy(np.concatenate(feats, axis=0)).pin_memory()
    host_views = [packed_host[int(o):int(o) + int(n)].numpy() for o, n in zip(offsets, n_frames)]
    packed_dev = packed_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        with torch.cuda.stream(stream):
            flush.zero_()                                               # L2 flush between iterations
        dec.decode_batch_device(packed_dev.data_ptr(), offsets, n_frames, want_results=False)

    # ---- value: features resident in HBM, CUDA events on the decoder's stream ----------
    for _ in range(args.warmup):
        device_step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = dec.launch_count
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    launches = dec.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    stats = dec.stats(-1)
    frames_step = rows
    ms_max, frames_all = jdist.reduce_time_and_frames(ms, frames_step * args.steps, dev)

    # ---- e2e: public host-buffer API, pinned inputs, H2D + D2H inside the timed region ---
    res = dec.decode_batch(host_views)                                   # warm-up + results for the parity check
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = dec.decode_batch(host_views)
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag.set()
    sampler.join(timeout=2)
    e2e_ms_max, _ = jdist.reduce_time_and_frames(1e3 * e2e_s, 0, dev)
    n_ok = sum(1 for r in res if r.status > 0)
    h2d = rows * m.dim * 4
    d2h = args.utts * 32 + 20 * sum(max(r.status, 0) for r in res) + 4   # ResHdr per utterance + the word pool in use

    # ---- roofline: one more step with per-kernel CUDA-event timing ----------------------
    dec.profile(True)
    device_step()
    prof = dec.profile_read()
    dec.profile(False)
    pstats = dec.stats(-1)
    dims = {"S": 5, "D": m.dim, "n_gmm": m.n_gmm, "C": max(len(w) for w in m.weights),
            "gmm_launches": prof["k_gmm_scores"]["launches"]}
    abytes = algorithmic_bytes(pstats, dims, rows)
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    top = max(prof, key=lambda k: prof[k]["ms"])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    top_ms_launch = prof[top]["ms"] / max(prof[top]["launches"], 1)
    achieved = (abytes[top] / max(prof[top]["launches"], 1)) / (top_ms_launch * 1e-3) / 1e9 if top_ms_launch > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        traffic = tj.get(args.workload, {}).get(top)
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": abytes[top] / max(prof[top]["launches"], 1),
        "avg_launch_us": 1e3 * top_ms_launch,
        "kernel_share_of_step": {k: round(v["ms"] / total_ms, 4) for k, v in prof.items()},
        "kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "kernel_launches": {k: v["launches"] for k, v in prof.items()},
        "kernel_gbs": {k: (abytes[k] / (prof[k]["ms"] * 1e-3) / 1e9 if prof[k]["ms"] > 0 else 0.0) for k in prof},
    }

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload ------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import binding
        from oracle.binding import OraclePort, OracleRef
        binding.build(ref=True, port=True)
        use_ref = os.path.exists(binding.REF_SO)
        if use_ref:
            o = OracleRef(files, **kw)
        else:
            from juicer_b200 import _abi
            o = OraclePort(_abi.FlatTables(network.arrays(), network.init_state, models.arrays()), _abi.make_cfg(**kw))
        order = np.argsort(n_frames)[: args.cpu_sample_utts]             # the shortest utterances: bounded CPU time
        fr = 0
        sec = 0.0
        parity = True
        for u in order:
            r = o.decode(feats[int(u)])
            fr += int(n_frames[u]); sec += r.seconds
            g = res[int(u)]
            parity &= (r.status == g.status and r.labels == g.labels and r.times == g.times
                       and abs(r.score - g.score) <= 1e-4)
        cpu = {"value": fr / sec, "unit": "frames/s", "cores": 1,
               "kind": "reference" if use_ref else "port",
               "sample": f"{len(order)} shortest utterances of this batch ({fr} frames), single thread, "
                         f"{'unmodified reference objects (oracle/_ref, g++ -O2)' if use_ref else 'plain-C restatement (gcc -O2)'}"}

    if rank == 0:
        value = frames_all / (ms_max * 1e-3)
        e2e_value = frames_step * args.steps * world / (e2e_ms_max * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload]}", "decoder": kw,
                       "utterances_per_step_per_gpu": args.utts, "frames_per_step_per_gpu": frames_step,
                       "lanes": args.lanes, "parallelism": f"utterance-sharded x{world}",
                       "l2": "256 MiB device buffer written between iterations; per-step state "
                             f"{(network.c.n_arcs * 12 + network.c.n_states * 8) * args.lanes / 1e9:.1f} GB >> 126 MB L2"},
            "xrt": value / 100.0,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "timing": "host wall clock around jgpu_decode_batch (pinned inputs), max over ranks",
                    "utterances_with_result": n_ok},
            "gpu_launches": int(launches),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "work_per_frame": {k: v / max(stats["n_frames"], 1) for k, v in stats.items() if k != "n_frames"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["parity_vs_cpu_sample"] = bool(parity)
        emit(line)
    dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
