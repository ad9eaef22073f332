"""TEST INFRASTRUCTURE ONLY — ctypes bindings for the two CPU checkers.

* ``OracleRef``  : oracle/_ref/liboracle_ref.so — the UNMODIFIED reference objects
                   (WFSTDecoderLite, HTKFlatModels, WFSTNetwork) behind oracle/ref_driver.cpp.
* ``OraclePort`` : oracle/liboracle.so — the plain-C restatement (oracle/juicer_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (juicer_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Dict, List, Optional

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "liboracle_ref.so")
REF_O3_SO = os.path.join(HERE, "_ref", "liboracle_ref_o3.so")
HARNESS_SO = os.path.join(HERE, "_ref", "liboracle_harness.so")     # the reference's own DecoderBatchTest harness (make harness)       # speed-only build (-O3, AVX2 + FMA), never used for parity
PORT_SO = os.path.join(HERE, "liboracle.so")
DROPIN_BIN = os.path.join(HERE, "_ref", "dropin_test")
LOG_ZERO = -np.finfo(np.float32).max


def build(ref: bool = True, port: bool = True) -> None:
    """Compile the checkers.  The reference build needs /root/reference (absent on the GPU
    box, where the prebuilt oracle/_ref/liboracle_ref.so that travelled with the snapshot is used)."""
    if port:
        subprocess.check_call(["make", "-s", "-C", HERE, "port"])
    if ref and os.path.isdir(os.environ.get("JUICER_REF", "/root/reference")):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        subprocess.check_call(["make", "-s", "-C", HERE, "ref_o3"])
        subprocess.check_call(["make", "-s", "-C", HERE, "harness"])
        if os.path.exists(os.path.join(HERE, "..", "juicer_b200", "libjuicer_b200.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin"])      # C++ adapter behind Juicer::IDecoder


class Word(C.Structure):
    _fields_ = [("label", C.c_int), ("time", C.c_int), ("score", C.c_float), ("ac", C.c_float), ("lm", C.c_float)]


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class DecodeResult:
    """status: number of words, -1 = no token reached a final state (reference returns NULL),
    -2 = surviving final token carries no word label (inactive DecHyp)."""

    def __init__(self, status: int, words: List[Dict], totals: np.ndarray,
                 frame_cnt: Optional[np.ndarray], frame_best: Optional[np.ndarray], seconds: float):
        self.status = status
        self.words = words
        self.score, self.ac, self.lm = (float(np.float32(x)) for x in totals)
        self.totals = totals
        self.frame_cnt = frame_cnt
        self.frame_best = frame_best
        self.seconds = seconds

    @property
    def labels(self) -> List[int]:
        return [w["label"] for w in self.words]

    @property
    def times(self) -> List[int]:
        return [w["time"] for w in self.words]

    def __repr__(self) -> str:
        return (f"DecodeResult(status={self.status}, labels={self.labels}, times={self.times}, "
                f"score={self.score!r}, ac={self.ac!r}, lm={self.lm!r})")


def _words_from(buf, n: int) -> List[Dict]:
    return [dict(label=buf[i].label, time=buf[i].time, score=float(np.float32(buf[i].score)),
                 ac=float(np.float32(buf[i].ac)), lm=float(np.float32(buf[i].lm))) for i in range(max(n, 0))]


class OracleRef:
    def __init__(self, files: Dict[str, str], *, main_beam: float, start_beam: float = 0.0,
                 end_beam: float = 0.0, word_beam: float = 0.0, max_hyps: int = 0,
                 lm_scale: float = 1.0, ins_penalty: float = 0.0, block_size: int = 5,
                 remove_tee: bool = False, so_path: Optional[str] = None):
        """files["jmbi"] (HTKFlatModels::readBinary) or, when only files["mmf"] is given, the MMF text form
        (HTKFlatModels::Load(mmf, remove_tee), on top of oracle/shim/htkparse_rd.cpp)."""
        so_path = so_path or REF_SO
        if not os.path.exists(so_path):
            raise RuntimeError(f"{so_path} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(so_path)
        self.lib.oref_create.restype = C.c_void_p
        self.lib.oref_create.argtypes = [C.c_char_p] * 4 + [C.c_float] * 6 + [C.c_int, C.c_int]
        self.lib.oref_decode.restype = C.c_int
        self.lib.oref_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.oref_gmm_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.oref_dims.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.oref_dump_models.argtypes = [C.c_void_p] * 11
        self.lib.oref_dump_net.argtypes = [C.c_void_p] * 8
        self.lib.oref_destroy.argtypes = [C.c_void_p]
        if "jmbi" in files:
            self.h = self.lib.oref_create(files["jmbi"].encode(), files["fsm"].encode(), files["insyms"].encode(),
                                          files["outsyms"].encode(), lm_scale, ins_penalty, start_beam, main_beam,
                                          end_beam, word_beam, max_hyps, block_size)
        else:
            self.lib.oref_create_mmf.restype = C.c_void_p
            self.lib.oref_create_mmf.argtypes = [C.c_char_p, C.c_int] + [C.c_char_p] * 3 + [C.c_float] * 6 + [C.c_int, C.c_int]
            self.h = self.lib.oref_create_mmf(files["mmf"].encode(), int(remove_tee), files["fsm"].encode(),
                                              files["insyms"].encode(), files["outsyms"].encode(), lm_scale, ins_penalty,
                                              start_beam, main_beam, end_beam, word_beam, max_hyps, block_size)
        d = np.zeros(16, dtype=np.int32)
        self.lib.oref_dims(self.h, _fp(d))
        (self.dim, self.n_gmm, self.n_hmm, self.n_tmat, self.max_states, self.max_comps,
         self.n_states, self.n_arcs, self.init_state) = (int(x) for x in d[:9])

    def close(self) -> None:
        if self.h:
            self.lib.oref_destroy(self.h)
            self.h = None

    def set_decoder(self, *, main_beam: float, start_beam: float = 0.0, end_beam: float = 0.0,
                    word_beam: float = 0.0, max_hyps: int = 0) -> None:
        """A new WFSTDecoderLite with other pruning settings on the same network and models."""
        self.lib.oref_set_decoder.argtypes = [C.c_void_p] + [C.c_float] * 4 + [C.c_int]
        self.lib.oref_set_decoder(self.h, start_beam, main_beam, end_beam, word_beam, max_hyps)

    def decode(self, feats: np.ndarray, counters: bool = False, max_words: int = 4096) -> DecodeResult:
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        T = feats.shape[0]
        words = (Word * max_words)()
        totals = np.zeros(3, dtype=np.float32)
        cnt = np.zeros((T, 6), dtype=np.int32) if counters else None
        best = np.zeros(T, dtype=np.float32) if counters else None
        sec = C.c_double(0.0)
        n = self.lib.oref_decode(self.h, _fp(feats), T, C.byref(words), max_words, _fp(totals),
                                 _fp(cnt) if counters else None, _fp(best) if counters else None, C.byref(sec))
        return DecodeResult(n, _words_from(words, min(n, max_words)), totals, cnt, best, sec.value)

    def gmm_scores(self, feats: np.ndarray) -> np.ndarray:
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        out = np.zeros((feats.shape[0], self.n_gmm), dtype=np.float32)
        self.lib.oref_gmm_scores(self.h, _fp(feats), feats.shape[0], _fp(out))
        return out

    def decode_partials(self, feats: np.ndarray, every: int):
        """Streaming partial results of the reference (tracePartialPath, src/WFSTDecoderLite.cpp:822-871) asked for
        after every `every`-th frame: [(frame, labels, word-end frames) or (frame, None, None) when not traceable]."""
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        T = feats.shape[0]
        max_tr = T // max(every, 1) + 1
        tf = np.zeros(max_tr, dtype=np.int32); tl = np.zeros(max_tr, dtype=np.int32)
        cap = max_tr * 512
        fl = np.zeros(cap, dtype=np.int32); ff = np.zeros(cap, dtype=np.int32)
        self.lib.oref_decode_partial.restype = C.c_int
        self.lib.oref_decode_partial.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int]
        n = self.lib.oref_decode_partial(self.h, _fp(feats), T, every, max_tr, _fp(tf), _fp(tl), _fp(fl), _fp(ff), cap)
        out, off = [], 0
        for k in range(n):
            if tl[k] < 0:
                out.append((int(tf[k]), None, None))
            else:
                out.append((int(tf[k]), fl[off:off + tl[k]].tolist(), ff[off:off + tl[k]].tolist()))
                off += int(tl[k])
        return out

    def dump_models(self) -> Dict[str, np.ndarray]:
        H, S, G, Cc, D = self.n_hmm, self.max_states, self.n_gmm, self.max_comps, self.dim
        r = dict(hmm_nstates=np.zeros(H, np.int32), hmm_gmm=np.zeros((H, S), np.int32),
                 hmm_tmat=np.zeros(H, np.int32), hmm_tee=np.zeros(H, np.float32),
                 trP=np.zeros((H, S, S), np.float32), se=np.zeros((H, S, 2), np.int32),
                 gmm_ncomp=np.zeros(G, np.int32), dets=np.zeros((G, Cc), np.float32),
                 means=np.zeros((G, Cc, D), np.float32), ivars=np.zeros((G, Cc, D), np.float32))
        self.lib.oref_dump_models(self.h, *[_fp(r[k]) for k in ("hmm_nstates", "hmm_gmm", "hmm_tmat", "hmm_tee",
                                                                "trP", "se", "gmm_ncomp", "dets", "means", "ivars")])
        return r

    def write_jwnt(self, path: str) -> None:
        """WFSTNetwork::writeBinary of the network as loaded (src/WFSTNetwork.cpp:1106-1226)."""
        self.lib.oref_write_jwnt.argtypes = [C.c_void_p, C.c_char_p]
        self.lib.oref_write_jwnt(self.h, path.encode())

    def dump_net(self) -> Dict[str, np.ndarray]:
        A, S = self.n_arcs, self.n_states
        r = dict(arc_to=np.zeros(A, np.int32), arc_w=np.zeros(A, np.float32), arc_in=np.zeros(A, np.int32),
                 arc_out=np.zeros(A, np.int32), st_first=np.zeros(S, np.int32), st_n=np.zeros(S, np.int32),
                 st_final=np.zeros(S, np.float32))
        self.lib.oref_dump_net(self.h, *[_fp(r[k]) for k in ("arc_to", "arc_w", "arc_in", "arc_out",
                                                             "st_first", "st_n", "st_final")])
        return r


class RefModels:
    """Acoustic models loaded by the reference alone: HTKFlatModels::Load (MMF text, parsed by
    oracle/shim/htkparse_rd.cpp, everything after the parse is the reference's own code) or ::readBinary (JMBI)."""

    def __init__(self, path: str, *, remove_tee: bool = False, block_size: int = 5):
        if not os.path.exists(REF_SO):
            raise RuntimeError(f"{REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = lib = C.CDLL(REF_SO)
        lib.oref_models_from_mmf.restype = C.c_void_p
        lib.oref_models_from_mmf.argtypes = [C.c_char_p, C.c_int, C.c_int]
        lib.oref_models_from_jmbi.restype = C.c_void_p
        lib.oref_models_from_jmbi.argtypes = [C.c_char_p, C.c_int]
        lib.oref_model_dims.argtypes = [C.c_void_p, C.c_void_p]
        lib.oref_dump_models.argtypes = [C.c_void_p] * 11
        lib.oref_write_models.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        lib.oref_gmm_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        lib.oref_destroy.argtypes = [C.c_void_p]
        with open(path, "rb") as f:
            binary = f.read(4) == b"JMBI"
        self.h = lib.oref_models_from_jmbi(path.encode(), block_size) if binary else \
            lib.oref_models_from_mmf(path.encode(), int(remove_tee), block_size)
        d = np.zeros(8, dtype=np.int32)
        lib.oref_model_dims(self.h, _fp(d))
        self.dim, self.n_gmm, self.n_hmm, self.n_tmat, self.max_states, self.max_comps = (int(x) for x in d[:6])

    dump_models = OracleRef.dump_models
    gmm_scores = OracleRef.gmm_scores

    def write(self, path: str, binary: bool) -> None:
        """HTKModels::output (src/HTKModels.cpp:993-1109): JMBI when binary, else the reference's MMF text."""
        self.lib.oref_write_models(self.h, path.encode(), int(binary))

    def close(self) -> None:
        if self.h:
            self.lib.oref_destroy(self.h)
            self.h = None


class RefJwnt:
    """A network read by the reference's own WFSTNetwork::readBinary (JWNT), tables through public getters."""

    def __init__(self, path: str, lm_scale: float = 1.0, ins_penalty: float = 0.0):
        self.lib = C.CDLL(REF_SO)
        self.lib.oref_net_from_jwnt.restype = C.c_void_p
        self.lib.oref_net_from_jwnt.argtypes = [C.c_char_p, C.c_float, C.c_float]
        self.lib.oref_net_dims.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.oref_dump_net.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        self.lib.oref_destroy.argtypes = [C.c_void_p]
        self.h = self.lib.oref_net_from_jwnt(path.encode(), lm_scale, ins_penalty)
        d = np.zeros(3, np.int32)
        self.lib.oref_net_dims(self.h, _fp(d))
        self.n_states, self.n_arcs, self.init_state = (int(x) for x in d)

    def dump_net(self) -> Dict[str, np.ndarray]:
        A, S = self.n_arcs, self.n_states
        r = dict(arc_to=np.zeros(A, np.int32), arc_w=np.zeros(A, np.float32), arc_in=np.zeros(A, np.int32),
                 arc_out=np.zeros(A, np.int32), st_first=np.zeros(S, np.int32), st_n=np.zeros(S, np.int32),
                 st_final=np.zeros(S, np.float32))
        self.lib.oref_dump_net(self.h, *[_fp(r[k]) for k in ("arc_to", "arc_w", "arc_in", "arc_out",
                                                             "st_first", "st_n", "st_final")])
        return r

    def close(self) -> None:
        if self.h:
            self.lib.oref_destroy(self.h)
            self.h = None


class OraclePort:
    """Plain-C restatement (oracle/juicer_oracle.c) on the flat tables of include/juicer_b200.h."""

    def __init__(self, tables, cfg):
        from juicer_b200 import _abi
        if not os.path.exists(PORT_SO):
            build(ref=False, port=True)
        self._abi = _abi
        self.lib = C.CDLL(PORT_SO)
        self.lib.jor_create.restype = C.c_void_p
        self.lib.jor_create.argtypes = [C.c_void_p] * 4
        self.lib.jor_decode.restype = C.c_int
        self.lib.jor_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p]
        self.lib.jor_gmm_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        self.lib.jor_stats.argtypes = [C.c_void_p, C.c_void_p]
        self.lib.jor_destroy.argtypes = [C.c_void_p]
        self.tables, self.cfg = tables, cfg
        self.h = self.lib.jor_create(C.byref(tables.net), C.byref(tables.hmm), C.byref(tables.gmm), C.byref(cfg))
        if not self.h:
            raise RuntimeError("jor_create failed (max_states > 8?)")

    def close(self) -> None:
        if self.h:
            self.lib.jor_destroy(self.h)
            self.h = None

    def decode(self, feats: np.ndarray, counters: bool = False, max_words: int = 4096) -> DecodeResult:
        abi = self._abi
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        T = feats.shape[0]
        words = (abi.JgpuWord * max_words)()
        res = abi.JgpuResult(0, 0, 0.0, 0.0, 0.0, max_words, C.cast(words, C.POINTER(abi.JgpuWord)))
        cnt = np.zeros((T, 6), dtype=np.int32) if counters else None
        best = np.zeros(T, dtype=np.float32) if counters else None
        sec = C.c_double(0.0)
        st = self.lib.jor_decode(self.h, _fp(feats), T, C.byref(res), _fp(cnt) if counters else None,
                                 _fp(best) if counters else None, C.byref(sec))
        totals = np.asarray([res.score, res.ac, res.lm], dtype=np.float32)
        return DecodeResult(st, abi.words_to_list(res), totals, cnt, best, sec.value)

    def gmm_scores(self, feats: np.ndarray) -> np.ndarray:
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        out = np.zeros((feats.shape[0], self.tables.n_gmm), dtype=np.float32)
        self.lib.jor_gmm_scores(self.h, _fp(feats), feats.shape[0], _fp(out))
        return out

    def stats(self) -> Dict[str, int]:
        s = self._abi.JgpuStats()
        self.lib.jor_stats(self.h, C.byref(s))
        return s.as_dict()


REF_FORMATS = {"verbose": 0, "trans": 1, "ref": 2, "mlf": 3, "xmlf": 4}      # DBTOutputFormat, src/DecoderBatchTest.h:29-38


def ref_harness_run(files: Dict[str, str], lexicon: str, list_file: str, out_file: str, fmt: str, *, main_beam: float,
                    start_beam: float = 0.0, end_beam: float = 0.0, word_beam: float = 0.0, max_hyps: int = 0,
                    sent_start: str = "", sent_end: str = "", remove_sil: bool = False, frames_per_sec: int = 100) -> str:
    """DecoderBatchTest::run of the UNMODIFIED reference (oracle/harness_driver.cpp) over a list of (extended) HTK feature file
    names; returns the text it wrote.  Runs in a child process: the reference's error() exits."""
    import subprocess as sp
    import sys
    code = (
        "import ctypes as C, sys\n"
        f"lib = C.CDLL({HARNESS_SO!r})\n"
        "lib.oref_harness_run.argtypes = [C.c_char_p] * 7 + [C.c_int] + [C.c_float] * 4 + [C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int]\n"
        "a = sys.argv[1:]\n"
        "rc = lib.oref_harness_run(*[x.encode() for x in a[:7]], int(a[7]), *[float(x) for x in a[8:12]], int(a[12]), a[13].encode(), a[14].encode(), int(a[15]), int(a[16]))\n"
        "sys.exit(rc)\n")
    args = [files["jmbi"], files["fsm"], files["insyms"], files["outsyms"], lexicon, list_file, out_file, str(REF_FORMATS[fmt]),
            str(start_beam), str(main_beam), str(end_beam), str(word_beam), str(max_hyps), sent_start, sent_end,
            str(int(remove_sil)), str(frames_per_sec)]
    p = sp.run([sys.executable, "-c", code] + args, capture_output=True, text=True, timeout=600)
    if p.returncode != 0:
        raise RuntimeError(f"reference harness failed ({p.returncode}): {p.stderr[-2000:]}")
    with open(out_file) as f:
        return f.read()
