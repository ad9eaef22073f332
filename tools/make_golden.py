"""Generates tests/golden/* by running the UNMODIFIED reference (oracle/_ref, built from
/root/reference by `make -C oracle ref`) on small seeded fixtures.

Run in the build container (needs /root/reference):   python tools/make_golden.py
Each fixture directory holds the exact input files (JMBI, FSM text, symbol tables), the
feature matrices, and expected.npz with the reference's outputs as raw float32 bit
patterns: word records, totals, per-frame work counters, bestEmitScore per frame and a
block of HTKFlatModels::calcOutput values.  The reference ships no golden vectors of its
own; these pin the oracle port and the CUDA path to the reference.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juicer_b200 import synth                      # noqa: E402
from oracle.binding import OracleRef, build        # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {                                           # name -> (config, n_utts, min_frames, truncate_last)
    "c1": ("c1", 2, 95, False),
    "c1h": ("c1h", 2, 95, False),
    "tee": ("tee", 2, 80, False),
    "mixed": ("mixed", 3, 150, False),
    "c2mini": ("c2mini", 2, 120, True),
    "nolabel": ("nolabel", 1, 40, False),
    "ties": ("ties", 3, 120, False),
}


def main() -> None:
    build(ref=True, port=False)
    only = set(sys.argv[1:])                        # `python tools/make_golden.py ties` regenerates one fixture
    for name, (cfg, n_utts, frames, trunc) in CASES.items():
        if only and name not in only:
            continue
        m, net, tee, kw = synth.named_config(cfg)
        d = os.path.join(GOLD, name)
        files = synth.make_fixture(name, d, m, net)
        ps = synth.PathSampler(net, m, tee_hmms=tee)
        rng = np.random.default_rng(1000 + len(name))
        o = OracleRef(files, **kw)
        o.write_jwnt(os.path.join(d, name + ".jwnt"))        # the same network through WFSTNetwork::writeBinary
        out = {}
        feats = []
        for u in range(n_utts):
            x, words = ps.sample(frames, rng)
            if trunc and u == n_utts - 1:
                x = x[: x.shape[0] // 2 + 3]         # ends mid-word: exercises "no final token" / partial paths
            feats.append(x)
        feats.append(np.zeros((0, m.dim), dtype=np.float32))   # empty utterance
        for u, x in enumerate(feats):
            r = o.decode(x, counters=True)
            out[f"x{u}"] = x
            out[f"status{u}"] = np.int32(r.status)
            out[f"totals{u}"] = r.totals.view(np.uint32)
            out[f"labels{u}"] = np.asarray(r.labels, dtype=np.int32)
            out[f"times{u}"] = np.asarray(r.times, dtype=np.int32)
            out[f"wscores{u}"] = np.asarray([[w["score"], w["ac"], w["lm"]] for w in r.words],
                                            dtype=np.float32).reshape(-1, 3).view(np.uint32)
            out[f"cnt{u}"] = r.frame_cnt[:, :5]
            out[f"best{u}"] = r.frame_best.view(np.uint32)
        g = o.gmm_scores(feats[0][:24])
        out["gmm_rows"] = np.int32(24)
        out["gmm"] = g.view(np.uint32)
        np.savez_compressed(os.path.join(d, "expected.npz"), **out)
        with open(os.path.join(d, "meta.json"), "w") as f:
            json.dump({"config": cfg, "decoder": kw, "n_utts": len(feats),
                       "generator": "tools/make_golden.py", "reference": "idiap/juicer @ c1d67eb, -O2, "
                       "-DOPT_FLATMODEL -DOPT_SINGLE_BEST -DPARTIAL_DECODING"}, f, indent=1)
        sz = sum(os.path.getsize(os.path.join(d, p)) for p in os.listdir(d))
        print(f"{name}: {len(feats)} utterances, statuses {[int(out[f'status{u}']) for u in range(len(feats))]}, {sz/1024:.0f} KiB")
        o.close()


if __name__ == "__main__":
    main()
