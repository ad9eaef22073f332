"""Shared helpers for the parity tests (test infrastructure: may use oracle/)."""
from __future__ import annotations

import json
import os
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["c1", "c1h", "tee", "mixed", "c2mini", "nolabel", "ties"]


def bits(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32)).view(np.uint32)


class Golden:
    def __init__(self, name: str):
        self.name = name
        self.dir = os.path.join(GOLD, name)
        with open(os.path.join(self.dir, "meta.json")) as f:
            self.meta = json.load(f)
        self.z = np.load(os.path.join(self.dir, "expected.npz"))
        self.n_utts = int(self.meta["n_utts"])
        self.kw: Dict = dict(self.meta["decoder"])
        self.files = {k: os.path.join(self.dir, f"{name}.{k}") for k in ("jmbi", "fsm", "insyms", "outsyms")}

    def feats(self, u: int) -> np.ndarray:
        return np.ascontiguousarray(self.z[f"x{u}"], dtype=np.float32)

    def check(self, u: int, res, frame_cnt=None, frame_best=None, what: str = "") -> None:
        """res: anything with status/labels/times/totals/words.  Bit-exact comparison."""
        z = self.z
        tag = f"{self.name}/utt{u} {what}"
        assert int(res.status) == int(z[f"status{u}"]), tag
        assert list(res.labels) == z[f"labels{u}"].tolist(), tag
        assert list(res.times) == z[f"times{u}"].tolist(), tag
        if int(res.status) > 0:
            assert np.array_equal(bits(res.totals), z[f"totals{u}"]), (tag, res.totals)
            ws = bits([[w["score"], w["ac"], w["lm"]] for w in res.words]).reshape(-1, 3)
            assert np.array_equal(ws, z[f"wscores{u}"]), tag
        if frame_cnt is not None:
            ref = z[f"cnt{u}"]
            if frame_cnt.shape[1] == 4:                 # GPU: insts, emit, end, endProcessed
                ref = ref[:, [0, 1, 2, 4]]
            else:
                frame_cnt = frame_cnt[:, :5]
            assert np.array_equal(frame_cnt, ref), tag
        if frame_best is not None:
            assert np.array_equal(bits(frame_best), z[f"best{u}"]), tag


MMF_CASES = ["tee", "mixed"]
MODEL_TABLES = ("hmm_nstates", "hmm_gmm", "hmm_tee", "trP", "se", "gmm_ncomp", "dets", "means", "ivars")


class _Prefixed:
    """npz view that prepends a key prefix (expected_mmf.npz holds one result set per load flag)."""

    def __init__(self, z, prefix: str):
        self.z, self.prefix = z, prefix

    def __getitem__(self, k: str):
        return self.z[self.prefix + k]


class GoldenMmf(Golden):
    """The MMF (HTK text) form of a golden fixture: <name>.mmf + expected_mmf.npz (tools/make_golden_mmf.py):
    the tables the reference holds after HTKFlatModels::Load(mmf, remove_tee) and its decode results."""

    def __init__(self, name: str, remove_tee: bool = False):
        super().__init__(name)
        self.remove_tee = bool(remove_tee)
        self.zx = self.z                                     # features live in expected.npz
        self.zm = np.load(os.path.join(self.dir, "expected_mmf.npz"))
        self.z = _Prefixed(self.zm, f"r{int(self.remove_tee)}_")
        self.files = dict(self.files, mmf=os.path.join(self.dir, name + ".mmf"))

    def feats(self, u: int) -> np.ndarray:
        return np.ascontiguousarray(self.zx[f"x{u}"], dtype=np.float32)

    def tables(self, prefix: str = None) -> Dict[str, np.ndarray]:
        prefix = prefix or f"r{int(self.remove_tee)}_tab_"
        return {k: self.zm[prefix + k] for k in MODEL_TABLES}


def same_model_tables(ref: Dict[str, np.ndarray], mine: Dict[str, np.ndarray], what: str = "") -> None:
    """Bit-exact comparison of the flat model tables (floats as uint32 bit patterns on either side)."""
    for k in MODEL_TABLES:
        x, y = np.asarray(ref[k]), np.asarray(mine[k])
        x = x.view(np.uint32) if x.dtype == np.float32 else x
        y = y.view(np.uint32) if y.dtype == np.float32 else y
        assert x.shape == y.shape, f"{what}: {k} shape {x.shape} vs {y.shape}"
        assert np.array_equal(x, y), f"{what}: {k}"


def flat_tables_from_mmf(files: Dict[str, str], remove_tee: bool = False, lm_scale: float = 1.0, ins_penalty: float = 0.0):
    from juicer_b200 import _abi, api
    net = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"], lm_scale, ins_penalty)
    models = api.HTKFlatModels.from_mmf(files["mmf"], remove_tee)
    tabs = _abi.FlatTables(net.arrays(), net.init_state, models.arrays())
    return tabs, net, models


def flat_tables_from_files(files: Dict[str, str], lm_scale: float = 1.0, ins_penalty: float = 0.0):
    """files -> product host loaders -> FlatTables (+ the loader objects that own the memory)."""
    from juicer_b200 import _abi, api
    net = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"], lm_scale, ins_penalty)
    models = api.HTKFlatModels(files["jmbi"])
    tabs = _abi.FlatTables(net.arrays(), net.init_state, models.arrays())
    return tabs, net, models


def same_result(a, b, what: str = "", exact: bool = True) -> None:
    assert a.status == b.status, (what, a, b)
    assert a.labels == b.labels and a.times == b.times, (what, a, b)
    if a.status > 0:
        if exact:
            assert np.array_equal(bits(a.totals), bits(b.totals)), (what, a, b)
            for x, y in zip(a.words, b.words):
                for k in ("score", "ac", "lm"):
                    assert np.float32(x[k]).tobytes() == np.float32(y[k]).tobytes(), (what, k, x, y)
        else:
            # north_star tolerance: path scores within 1e-4; acoustic totals (|x| ~ 1e4..1e5,
            # fp32 ulp >= 1e-3) relative 1e-6
            assert abs(a.score - b.score) <= 1e-4, (what, a, b)
            assert abs(a.lm - b.lm) <= 1e-4 * max(1.0, abs(a.lm) / 100.0), (what, a, b)
            assert abs(a.ac - b.ac) <= 1e-6 * abs(a.ac) + 1e-4, (what, a, b)
