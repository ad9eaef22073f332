#!/usr/bin/env python
"""bench.py — decoded frames/sec of the WFST token-passing decode path on B200.

    python bench.py --gpus N --steps K --warmup W                 (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K --warmup W  (reference CPU path)

A "step" decodes one batch of synthetic utterances (39-d MFCC-like frames sampled along
random accepted paths) on the BASELINE.json configs[2]-shaped network ("c3": 20k-word
trigram-shaped C.L.G, ~440k states / ~1.8M arcs, 6000 tied 16-mix GMMs, main beam 250).
One process per GPU; utterances are independent, so there is no data-path collective:

* N = 1           : the rank decodes its batch of --utts utterances (jgpu_decode_batch[_device]).
* N > 1, weak     : the batch is N x --utts utterances (rank r contributes the list seeded 1000 + r; every
                    rank holds all of them) behind ONE shared queue — a rank whose lanes run dry claims the next
                    utterance with an atomic fetch-add on a node-local shared-memory counter (jgpu_decode_queue:
                    whole-utterance work stealing, BASELINE configs[3]'s mechanism).  Work grows with N.
* --scaling strong: BASELINE configs[3] itself: --total-utts utterances (default 10000) in all, whatever N.
torch.distributed (NCCL) is used for barriers, the timing reduction (max over ranks) and the frame counts.

Prints ONE JSON line (rank 0).  `value` = frames/s with features resident in HBM, timed with
CUDA events on the decoder's stream; `e2e` = the same batches through the host-buffer entry point with
pinned HOST buffers (H2D feature copies and D2H result copies inside the timed region).
"""
from __future__ import annotations

import argparse
import glob
import hashlib
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
    "c3": "20k-word trigram-shaped C.L.G (~440k states / ~1.8M arcs), 4000 HMMs over 6000 tied 16-mix GMMs, "
          "main beam 250, utterances of 300-1000 frames",
    "c3p": "c3 with a prefix-tree lexicon at the hub (shared first phones, word label on the first unique arc, pushed "
           "unigram weights; ~444k states / ~1.8M arcs), main beam 250, utterances of 300-1000 frames",
    "c2": "1k-word bigram C.L.G (~42k states / ~51k arcs), 2000 triphone HMMs x 16-mix, main beam 200, "
          "utterances of 300-1000 frames",
    "c3s": "c3 topology at 1/8 scale (smoke runs)",
    "c3ps": "c3p topology at 1/8 scale (smoke runs)",
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
    ap.add_argument("--utts", type=int, default=512, help="utterances per step per GPU (weak scaling)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-utts", type=int, default=10000, help="utterances per step in all (strong scaling)")
    ap.add_argument("--lanes", type=int, default=256, help="utterances decoded in lock-step")
    ap.add_argument("--min-frames", type=int, default=300)
    ap.add_argument("--max-frames", type=int, default=1000)
    ap.add_argument("--cpu-sample-utts", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the c3p side-by-side measurement of the default run")
    ap.add_argument("--beam", type=float, default=0.0, help="override the workload's main beam (beam sweeps, BASELINE configs[4])")
    ap.add_argument("--max-hyps", type=int, default=-1, help="override the histogram-pruning limit (0 = off)")
    ap.add_argument("--start-beam", type=float, default=-1.0)
    ap.add_argument("--end-beam", type=float, default=-1.0, help="phone-end beam (> 0 takes the k_filter path)")
    ap.add_argument("--word-beam", type=float, default=-1.0)
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


def make_config(args, kw, world: int, net=None) -> Dict:
    """The `config` object: identical in both arms (the reference arm decodes a bounded sample of the same batch)."""
    strong = args.scaling == "strong"
    state_gb = (net.n_arcs * 12 + net.n_states * 8) * args.lanes / 1e9 if net is not None else 0.0
    return {"workload": f"{args.workload}: {WORKLOADS[args.workload]}", "decoder": kw,
            "l2": f"CUDA arm: 256 MiB device buffer written between iterations; per-step state {state_gb:.1f} GB >> 126 MB L2",
            "utterances_per_step": args.total_utts if strong else args.utts * world,
            "utterances_per_step_per_gpu": None if strong else args.utts,
            "utterance_frames": [args.min_frames, args.max_frames], "utterance_seed": "1000 + rank",
            "lanes": args.lanes,
            "parallelism": f"utterance-sharded x{world}" + (", one shared utterance queue (work stealing)" if world > 1 or strong else "")}


def csrc_sha() -> str:
    """Hash of the kernel sources: an ncu traffic capture is only quoted for the code it was taken on."""
    h = hashlib.sha256()
    for p in sorted(glob.glob(os.path.join(ROOT, "juicer_b200", "csrc", "*.cu*"))):
        h.update(os.path.basename(p).encode())
        h.update(open(p, "rb").read())
    return h.hexdigest()[:16]


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


def _ref_init(files, kw, kind):
    """Pool initializer: every worker process loads the network and models once."""
    global _REF_DEC
    sys.path.insert(0, ROOT)
    from oracle import binding
    from oracle.binding import OraclePort, OracleRef
    if kind == "reference":
        _REF_DEC = OracleRef(files, **kw)
    elif kind == "reference_o3":
        _REF_DEC = OracleRef(files, so_path=binding.REF_O3_SO, **kw)
    else:
        from juicer_b200 import _abi, api
        net = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
        models = api.HTKFlatModels(files["jmbi"])
        _REF_DEC = OraclePort(_abi.FlatTables(net.arrays(), net.init_state, models.arrays()), _abi.make_cfg(**kw))
        _REF_DEC._keep = (net, models)


def _ref_worker(x):
    t0 = time.perf_counter()
    r = _REF_DEC.decode(x, counters=True)
    dt = time.perf_counter() - t0
    cnt = r.frame_cnt
    return x.shape[0], dt, int(cnt[:, 1].sum()) if cnt is not None else 0, int(cnt[:, 0].sum()) if cnt is not None else 0


def _time_cpu_pool(files, kw, kind, feats, cores, warmup, steps):
    """`steps` timed repeats of one pass over `feats` (one utterance per core): per-repeat frames/s."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    per_step, emit, insts, frames = [], 0, 0, 0
    with ctx.Pool(cores, initializer=_ref_init, initargs=(files, kw, kind)) as pool:
        pool.map(_ref_worker, [f[:2] for f in feats], chunksize=1)        # force every worker to finish loading
        for step in range(warmup + steps):
            t0 = time.perf_counter()
            res = pool.map(_ref_worker, feats, chunksize=1)
            dt = time.perf_counter() - t0                                  # wall time of the step (all cores busy)
            if step >= warmup:
                per_step.append(sum(r[0] for r in res) / dt)
                emit += sum(r[2] for r in res); insts += sum(r[3] for r in res); frames += sum(r[0] for r in res)
    return per_step, emit / max(frames, 1), insts / max(frames, 1), frames


def run_reference(args) -> None:
    """CPU baseline = the reference's own WFSTDecoderLite/HTKFlatModels objects (oracle/_ref, kind
    "reference") when the prebuilt library is present, else the plain-C restatement (kind "port").
    The reference is single-threaded; "all the host threads it can use" = one independent process
    per core over a disjoint split of the utterances.  The utterances are the
    shortest ones of the batch the CUDA arm decodes (same seed), one per core per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import binding
    binding.build(ref=True, port=True)
    kind = "reference" if os.path.exists(binding.REF_SO) else "port"
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    m, net, tee, kw, files = build_fixture(args.workload, args.workdir, 0)
    batch = sample_utterances(net, m, tee, args.utts, args.min_frames, args.max_frames, seed=1000)   # rank 0's batch of the CUDA arm
    order = np.argsort([f.shape[0] for f in batch], kind="stable")[:cores]
    feats = [batch[int(u)] for u in order]
    per_step, emit, insts, frames = _time_cpu_pool(files, kw, kind, feats, cores, args.warmup, args.steps)
    value = float(np.mean(per_step))
    flags = "unmodified reference objects, g++ -O2 (the parity build)" if kind == "reference" else "plain-C restatement, gcc -O2"
    sample = (f"the {len(feats)} shortest utterances of rank 0's batch of the CUDA arm ({min(f.shape[0] for f in feats)}-"
              f"{max(f.shape[0] for f in feats)} frames), one per host core per step; {flags}")
    cpu = {"value": value, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
           "repeats": len(per_step), "best": float(np.max(per_step)), "worst": float(np.min(per_step)),
           "avgActiveEmitHyps": emit, "avgActiveModels": insts}
    best_effort = None
    if kind == "reference" and os.path.exists(binding.REF_O3_SO):
        ps3, emit3, _insts3, _ = _time_cpu_pool(files, kw, "reference_o3", feats, cores, min(args.warmup, 1), args.steps)
        best_effort = {"value": float(np.mean(ps3)), "best": float(np.max(ps3)), "worst": float(np.min(ps3)), "unit": "frames/s",
                       "cores": cores, "build": "the same sources, g++ -O3 -march=x86-64-v3 -ffp-contract=fast (speed only: GMM values "
                                                "differ from the parity build by <= 2 ulp in ~1 % of cases)",
                       "avgActiveEmitHyps": emit3}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (frames / max(len(per_step), 1)) / value,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": make_config(args, kw, world, net),
        "cpu_baseline": cpu,
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "xrt": value / 100.0,
        "note": "one host, whatever N: the CPU arm does not scale with the number of GPUs",
    }
    if best_effort:
        line["best_effort_cpu"] = best_effort
    emit_line(line)


# ---------------------------------------------------------------------------------------
# roofline bookkeeping
# ---------------------------------------------------------------------------------------
def algorithmic_bytes(stats: Dict[str, int], dims: Dict[str, int], n_rows: int) -> Dict[str, Dict]:
    """Algorithmic bytes moved by each kernel over one whole step, from the decoder's own work
    counters and the device record sizes of DESIGN.md ("Kernels"): instance record 16 B, token 16 B,
    arrival record 32 B, state row 16 B, arc 16 B, slotmap entry 4 B, word-boundary record 32 B.
    Minimal traffic: every record counted once.  `level` says where those bytes live: "hbm" = per-lane
    state streamed from DRAM once per step; "l2" = the static network rows and the arrival records just
    written, which are served by the 126 MB L2 (their rate is NOT an HBM rate)."""
    A = stats["total_active_models"]          # instances walked by the internal phase (sum over frames)
    H = stats["total_active_emit_hyps"]       # live emitting tokens after the internal phase
    Ea = stats["total_active_end_hyps"]       # live exit tokens = arrival records written by k_internal
    E = stats["total_proc_end_hyps"]          # exit tokens passing the end/word beam = first-round arrivals
    X = stats["total_arcs_expanded"]          # out-arcs of the states committed (hub rows included)
    W = stats["total_entry_writes"]           # distinct destination arcs whose entry token was written
    P = stats["total_paths"]                  # word-boundary records appended
    D, G, M = dims["D"], dims["n_gmm"], dims["C"]
    pairs = stats["total_gmm_evals"]          # (GMM, frame) scores computed
    return {
        # instance record + entry token read per instance; emitting tokens read + written, one acoustic score
        # per live token; one arrival record per live exit token
        "k_internal": {"bytes": A * (16 + 16) + H * (16 + 16 + 4) + Ea * 32, "level": "hbm"},
        # (only with an end / word beam) arrival records re-read and filtered
        "k_filter": {"bytes": Ea * 16, "level": "l2"},
        # expansion round 0: arrival record + state row per record, word-boundary records
        "k_expand": {"bytes": E * (32 + 16) + P * 32, "level": "l2"},
        "k_expand_r1": {"bytes": 0.0, "level": "l2"}, "k_expand_r2": {"bytes": 0.0, "level": "l2"},
        # commit: arrival record + state row per record, arc + slotmap entry per arc walked, entry token +
        # instance record + slotmap entry per entry written (rows of hub-like states are walked by
        # k_commit_huge; their bytes are counted here)
        "k_commit": {"bytes": E * (32 + 16) + X * (16 + 4) + W * (16 + 16 + 4), "level": "l2"},
        "k_commit_huge": {"bytes": 0.0, "level": "l2"},
        "k_boundary": {"bytes": 0.0, "level": "l2"},
        # scores out, features in, one pass over the (L2-resident) parameters per launch
        "k_gmm_scores": {"bytes": pairs * 4 + n_rows * D * 4 + dims["gmm_launches"] * G * M * (2 * D + 1) * 4, "level": "l2"},
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


def emit_line(line: Dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


class Batch:
    """The utterances of a step: host copies (pinned, packed), device copy, frame counts."""

    def __init__(self, feats, dev, torch):
        self.feats = feats
        self.n_frames = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
        self.offsets = np.concatenate([[0], np.cumsum(self.n_frames)[:-1]]).astype(np.int64)
        self.rows = int(self.n_frames.sum())
        self.packed_host = torch.from_numpy(np.concatenate(feats, axis=0)).pin_memory()
        self.host_views = [self.packed_host[int(o):int(o) + int(n)].numpy() for o, n in zip(self.offsets, self.n_frames)]
        self.packed_dev = self.packed_host.to(dev)
        self.order = np.argsort(-self.n_frames, kind="stable").astype(np.int32)     # longest first


def measure(args, dec, batch, stream, flush, jdist, dist, world, dev, torch, use_queue):
    """value (device-resident features, CUDA events) and e2e (host buffers, wall clock) of `batch`."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    claimed_log: List[int] = []
    busy_log: List[float] = []

    def device_step():
        with torch.cuda.stream(stream):
            flush.zero_()                                               # L2 flush between iterations
        if use_queue:
            q = jdist.SharedQueue.collective("bench")                  # fresh counter, two barriers
            mine, busy = dec.decode_queue_device(q, batch.packed_dev.data_ptr(), batch.offsets, batch.n_frames, batch.order)
            claimed_log.append(len(mine)); busy_log.append(busy)
            if world > 1:
                dist.barrier()
            q.close()
        else:
            dec.decode_batch_device(batch.packed_dev.data_ptr(), batch.offsets, batch.n_frames, want_results=False)

    for _ in range(args.warmup):
        device_step()
    barrier()
    del claimed_log[:], busy_log[:]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = dec.launch_count
    ev0.record(stream)
    for _ in range(args.steps):
        device_step()
    ev1.record(stream)
    barrier()
    launches = dec.launch_count - launches0
    ms = ev0.elapsed_time(ev1)
    stats = dec.stats(-1)                                               # of this rank's last step
    ms_max, _ = jdist.reduce_time_and_frames(ms, 0, dev)
    frames_all = batch.rows * args.steps                                # the whole batch is decoded once per step

    # ---- e2e: public host-buffer API, pinned inputs, H2D + D2H inside the timed region ---
    def host_step():
        if use_queue:
            q = jdist.SharedQueue.collective("bench")
            mine, _busy = dec.decode_queue(q, batch.host_views, batch.order)
            if world > 1:
                dist.barrier()
            q.close()
            return mine
        return dict(enumerate(dec.decode_batch(batch.host_views)))

    res = host_step()                                                   # warm-up + results for the parity check
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = host_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_ms_max, _ = jdist.reduce_time_and_frames(1e3 * e2e_s, 0, dev)
    mine_rows = int(sum(batch.n_frames[u] for u in res))
    h2d = mine_rows * batch.feats[0].shape[1] * 4
    d2h = len(res) * 32 + 20 * sum(max(r.status, 0) for r in res.values()) + 4   # ResHdr per utterance + the word pool in use
    n_ok = sum(1 for r in res.values() if r.status > 0)
    out = {"ms_max": ms_max, "frames_all": frames_all, "launches": launches, "stats": stats,
           "e2e_ms_max": e2e_ms_max, "h2d": h2d, "d2h": d2h, "n_ok": n_ok, "n_mine": len(res), "res": res}
    if use_queue:
        # per-rank view of the shared queue in the timed device-resident steps
        cl = torch.tensor([float(np.mean(claimed_log)), float(np.mean(busy_log))], dtype=torch.float64, device=dev)
        if world > 1:
            parts = [torch.zeros_like(cl) for _ in range(world)]
            dist.all_gather(parts, cl)
        else:
            parts = [cl]
        out["queue"] = {"claimed_per_rank_per_step": [float(p[0]) for p in parts],
                        "busy_ms_per_rank_per_step": [float(p[1]) for p in parts]}
    return out


def main() -> None:
    args = parse_args()
    claim_stdout()
    if args.beam > 0:
        _OVERRIDES["main_beam"] = args.beam
    if args.max_hyps >= 0:
        _OVERRIDES["max_hyps"] = args.max_hyps
    for k, v in (("start_beam", args.start_beam), ("end_beam", args.end_beam), ("word_beam", args.word_beam)):
        if v >= 0:
            _OVERRIDES[k] = v
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
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # stdout carries exactly one JSON line
        jdist.init_process_group("nccl")
        import torch.distributed as dist
    api.load_library()
    strong = args.scaling == "strong"
    use_queue = world > 1 or strong

    def make_decoder(workload):
        m, net, tee, kw, files = build_fixture(workload, args.workdir, rank)
        network = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
        models = api.HTKFlatModels(files["jmbi"])
        dec = api.WFSTDecoderLite(network, models, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                                  kw.get("word_beam", 0.0), kw.get("max_hyps", 0), n_lanes=args.lanes, device=local_rank)
        return m, net, tee, kw, files, network, models, dec

    m, net, tee, kw, files, network, models, dec = make_decoder(args.workload)
    stream = torch.cuda.Stream(device=dev)
    dec.set_stream(stream.cuda_stream)

    # the batch of a step
    if strong:
        base = sample_utterances(net, m, tee, min(args.total_utts, 1024), args.min_frames, args.max_frames, seed=1000)
        feats = [base[i % len(base)] for i in range(args.total_utts)]     # 10k utterances = 1024 distinct ones, cycled
        own = feats[rank::world]
    else:
        lists = [sample_utterances(net, m, tee, args.utts, args.min_frames, args.max_frames, seed=1000 + r) for r in range(world)]
        feats = [f for lst in lists for f in lst]
        own = lists[rank]
    batch = Batch(feats, dev, torch)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    r = measure(args, dec, batch, stream, flush, jdist, dist, world, dev, torch, use_queue)
    sampler.stop_flag.set()
    sampler.join(timeout=2)

    # ---- roofline: one more pass over the rank's own utterances with per-kernel CUDA-event timing ----------
    own_batch = batch if len(own) == len(feats) else Batch(own[:512], dev, torch)
    dec.profile(True)
    with torch.cuda.stream(stream):
        flush.zero_()
    dec.decode_batch_device(own_batch.packed_dev.data_ptr(), own_batch.offsets, own_batch.n_frames, want_results=False)
    prof = dec.profile_read()
    dec.profile(False)
    pstats = dec.stats(-1)
    fp32_peak = dec.fp32_peak_tops()
    n_comp = max(len(w) for w in m.weights)
    dims = {"S": 5, "D": m.dim, "n_gmm": m.n_gmm, "C": n_comp, "gmm_launches": prof["k_gmm_scores"]["launches"]}
    ab = algorithmic_bytes(pstats, dims, own_batch.rows)
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    top = max((k for k in prof if ab[k]["level"] == "hbm"), key=lambda k: prof[k]["ms"])      # the dominant HBM-bound kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    n_top = max(prof[top]["launches"], 1)
    top_ms_launch = prof[top]["ms"] / n_top
    achieved = (ab[top]["bytes"] / n_top) / (top_ms_launch * 1e-3) / 1e9 if top_ms_launch > 0 else 0.0
    traffic, traffic_note = None, "no ncu capture committed for this workload"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = tj.get(args.workload, {})
        if ent.get("csrc_sha") == csrc_sha():
            traffic, traffic_note = ent.get(top), ent.get("_note")
        elif ent:
            traffic_note = (f"profiles/roofline_traffic.json was captured on kernel sources {ent.get('csrc_sha')}, this run is "
                            f"{csrc_sha()}: not quoted")
    except Exception:
        pass
    # the scorer: FP32 operations without FMA: sub, mul, mul, add per (Gaussian, dimension)
    gmm_ms = prof["k_gmm_scores"]["ms"]
    gmm_ops = pstats["total_gmm_evals"] * n_comp * (4.0 * m.dim)
    gmm_tops = gmm_ops / (gmm_ms * 1e-3) / 1e12 if gmm_ms > 0 else 0.0
    roofline = {
        "bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic, "traffic_note": traffic_note, "csrc_sha": csrc_sha(), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": ab[top]["bytes"] / n_top,
        "avg_launch_us": 1e3 * top_ms_launch,
        "kernel_share_of_step": {k: round(v["ms"] / total_ms, 4) for k, v in prof.items()},
        "kernel_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
        "kernel_launches": {k: v["launches"] for k, v in prof.items()},
        # rate of each kernel's algorithmic bytes and WHERE those bytes live: only "hbm" entries compare with hbm_gbs
        "kernel_algorithmic_gbs": {k: {"gbs": round(ab[k]["bytes"] / (prof[k]["ms"] * 1e-3) / 1e9, 1) if prof[k]["ms"] > 0 else 0.0,
                                       "level": ab[k]["level"]} for k in prof},
        "secondary": {"kernel": ("k_gmm_scores = k_gmm_lazy over the stamped (GMM, lane) pairs of every step (JUICER_B200_LAZY=1)"
                                 if os.environ.get("JUICER_B200_LAZY", "0") not in ("", "0")
                                 else "k_gmm_scores: every GMM for every (lane, frame), 16 frames ahead per launch"),
                      "bound": "fp32 issue, no FMA", "achieved": gmm_tops, "peak": fp32_peak, "unit": "T op/s",
                      "frac": gmm_tops / fp32_peak if fp32_peak else None,
                      "ops": "4 per (Gaussian, dimension): sub, mul, mul, add", "peak_source": "jgpu_ubench_fp32 in this run",
                      "gmm_frame_pairs_scored": pstats["total_gmm_evals"],
                      "fraction_of_dense": pstats["total_gmm_evals"] / max(own_batch.rows * m.n_gmm, 1)},
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
        order = np.argsort(batch.n_frames, kind="stable")[: args.cpu_sample_utts]   # the shortest utterances: bounded CPU time
        fr, sec, emit = 0, 0.0, 0
        parity = True
        for u in order:
            rr = o.decode(feats[int(u)], counters=True)
            fr += int(batch.n_frames[u]); sec += rr.seconds; emit += int(rr.frame_cnt[:, 1].sum())
            g = r["res"].get(int(u))
            parity &= (g is not None and rr.status == g.status and rr.labels == g.labels and rr.times == g.times
                       and abs(rr.score - g.score) <= 1e-4)
        cpu = {"value": fr / sec, "unit": "frames/s", "cores": 1,
               "kind": "reference" if use_ref else "port", "avgActiveEmitHyps": emit / max(fr, 1),
               "sample": f"{len(order)} shortest utterances of this batch ({fr} frames), single thread, "
                         f"{'unmodified reference objects (oracle/_ref, g++ -O2)' if use_ref else 'plain-C restatement (gcc -O2)'}"}

    # ---- side by side (N = 1): the same network with a prefix-tree lexicon at the hub ----
    side = None
    if world == 1 and not strong and args.workload == "c3" and not args.no_side:
        dec.close()
        dec = None
        m2, net2, tee2, kw2, files2, network2, models2, dec2 = make_decoder("c3p")
        dec2.set_stream(stream.cuda_stream)
        b2 = Batch(sample_utterances(net2, m2, tee2, args.utts, args.min_frames, args.max_frames, seed=1000), dev, torch)
        a2 = argparse.Namespace(**dict(vars(args), warmup=1, steps=2))
        r2 = measure(a2, dec2, b2, stream, flush, jdist, dist, world, dev, torch, False)
        nf2 = max(r2["stats"]["n_frames"], 1)
        side = {"c3p": {"workload": WORKLOADS["c3p"], "value": r2["frames_all"] / (r2["ms_max"] * 1e-3), "unit": "frames/s",
                        "e2e": b2.rows * a2.steps / (r2["e2e_ms_max"] * 1e-3), "steps": a2.steps, "warmup": a2.warmup,
                        "ms_per_step": r2["ms_max"] / a2.steps,
                        "work_per_frame": {k: v / nf2 for k, v in r2["stats"].items() if k != "n_frames"}}}
        dec2.close()

    if rank == 0:
        stats = r["stats"]
        value = r["frames_all"] / (r["ms_max"] * 1e-3)
        e2e_value = batch.rows * args.steps / (r["e2e_ms_max"] * 1e-3)
        cfg = make_config(args, kw, world, net)
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_max"] / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg,
            "xrt": value / 100.0, "frames_per_step": batch.rows,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": r["h2d"], "d2h_bytes_per_step": r["d2h"],
                    "bytes_are": "rank 0's share of the batch" if use_queue else "the whole batch",
                    "timing": "host wall clock around jgpu_decode_batch / jgpu_decode_queue (pinned inputs), max over ranks",
                    "utterances_with_result": r["n_ok"], "utterances_decoded_by_rank0": r["n_mine"]},
            "gpu_launches": int(r["launches"]),
            "clocks": sampler.summary(),
            "roofline": roofline,
            "work_per_frame": {k: v / max(stats["n_frames"], 1) for k, v in stats.items() if k != "n_frames"},
        }
        if "queue" in r:
            line["queue"] = r["queue"]
        if cpu is not None:
            line["cpu_baseline"] = cpu
            line["parity_vs_cpu_sample"] = bool(parity)
        if side is not None:
            line["side_by_side"] = side
        emit_line(line)
    if dec is not None:
        dec.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
