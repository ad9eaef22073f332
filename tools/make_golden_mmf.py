"""Generates the MMF (HTK text model definition) fixtures of tests/golden/{tee,mixed}:

  <case>.mmf          the fixture's model set written HTK-style by juicer_b200.synth.write_mmf (macros ~o ~v ~t ~s ~h)
  tee.refout.mmf      the same models as written by the reference's own HTKModels::output(fName, false)
                      (src/HTKModels.cpp:993-1048, %.4e precision), minus the <trP>/<SEIndex> debug blocks it
                      appends to every in-HMM transition matrix (:1893-1910), which its own parser cannot read
  expected_mmf.npz    for removeInitialToFinalTransitions in {0, 1}: the tables the reference holds after
                      HTKFlatModels::Load(mmf, flag) and its decode results (word records, totals, per-frame
                      counters, bestEmitScore as raw float32 bit patterns) on the fixture's utterances

Run in the build container (needs /root/reference):   python tools/make_golden_mmf.py
The reference side is oracle/_ref: the unmodified reference objects on top of oracle/shim/htkparse_rd.cpp
(bison/flex do not exist here, so the generated parser itself cannot be built; everything after the parse is the
reference's own compiled code)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import Golden                          # noqa: E402
from juicer_b200 import synth                       # noqa: E402
from oracle.binding import OracleRef, RefModels, build   # noqa: E402

CASES = {"tee": dict(upper=True), "mixed": dict(upper=False)}
TABLES = ("hmm_nstates", "hmm_gmm", "hmm_tee", "trP", "se", "gmm_ncomp", "dets", "means", "ivars")


def strip_debug_blocks(text: str) -> str:
    out, skip = [], False
    for ln in text.split("\n"):
        if ln.startswith("<trP>") or ln.startswith("<SEIndex>"):
            skip = True
            continue
        if skip:
            if ln.strip() == "":
                skip = False
            continue
        out.append(ln)
    return "\n".join(out)


def main() -> None:
    build(ref=True, port=False)
    for name, style in CASES.items():
        g = Golden(name)
        m, net, tee, kw = synth.named_config(g.meta["config"])
        mmf = os.path.join(g.dir, name + ".mmf")
        synth.write_mmf(m, mmf, **style)
        out = {}
        for rt in (0, 1):
            rm = RefModels(mmf, remove_tee=bool(rt))
            tabs = rm.dump_models()
            for k in TABLES:
                out[f"r{rt}_tab_{k}"] = tabs[k].view(np.uint32) if tabs[k].dtype == np.float32 else tabs[k]
            rm.close()
            files = dict(mmf=mmf, fsm=g.files["fsm"], insyms=g.files["insyms"], outsyms=g.files["outsyms"])
            o = OracleRef(files, remove_tee=bool(rt), **kw)
            for u in range(g.n_utts):
                r = o.decode(g.feats(u), counters=True)
                p = f"r{rt}_"
                out[p + f"status{u}"] = np.int32(r.status)
                out[p + f"totals{u}"] = r.totals.view(np.uint32)
                out[p + f"labels{u}"] = np.asarray(r.labels, dtype=np.int32)
                out[p + f"times{u}"] = np.asarray(r.times, dtype=np.int32)
                out[p + f"wscores{u}"] = np.asarray([[w["score"], w["ac"], w["lm"]] for w in r.words],
                                                    dtype=np.float32).reshape(-1, 3).view(np.uint32)
                out[p + f"cnt{u}"] = r.frame_cnt[:, :5]
                out[p + f"best{u}"] = r.frame_best.view(np.uint32)
            print(name, "removeTee", rt, "statuses", [int(out[f"r{rt}_status{u}"]) for u in range(g.n_utts)],
                  "labels0", out[f"r{rt}_labels0"].tolist())
            o.close()
        if name == "tee":
            rj = RefModels(g.files["jmbi"])
            ref_txt = os.path.join(g.dir, name + ".refout.mmf")
            rj.write(ref_txt, False)
            rj.close()
            with open(ref_txt) as f:
                t = strip_debug_blocks(f.read())
            with open(ref_txt, "w") as f:
                f.write(t)
            rm = RefModels(ref_txt)
            tabs = rm.dump_models()
            for k in TABLES:
                out[f"refout_tab_{k}"] = tabs[k].view(np.uint32) if tabs[k].dtype == np.float32 else tabs[k]
            rm.close()
        np.savez_compressed(os.path.join(g.dir, "expected_mmf.npz"), **out)
        print(name, {f: os.path.getsize(os.path.join(g.dir, f)) // 1024 for f in os.listdir(g.dir) if "mmf" in f}, "KiB")


if __name__ == "__main__":
    main()
