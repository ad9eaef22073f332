"""ctypes mirror of include/juicer_b200.h (POD structs only; no device code here)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List

import numpy as np

LOG_ZERO = float(-np.finfo(np.float32).max)

i32p = C.POINTER(C.c_int32)
f32p = C.POINTER(C.c_float)


class JgpuNet(C.Structure):
    _fields_ = [("n_states", C.c_int32), ("n_arcs", C.c_int32), ("init_state", C.c_int32),
                ("arc_to", i32p), ("arc_weight", f32p), ("arc_in", i32p), ("arc_out", i32p),
                ("state_first", i32p), ("state_narcs", i32p), ("state_final", f32p)]


class JgpuHmm(C.Structure):
    _fields_ = [("n_hmms", C.c_int32), ("max_states", C.c_int32), ("n_states", i32p), ("gmm", i32p),
                ("trp", f32p), ("se", i32p), ("tee", f32p)]


class JgpuGmm(C.Structure):
    _fields_ = [("n_gmms", C.c_int32), ("dim", C.c_int32), ("max_comps", C.c_int32), ("n_comps", i32p),
                ("dets", f32p), ("means", f32p), ("ivars", f32p)]


class JgpuCfg(C.Structure):
    _fields_ = [("start_beam", C.c_float), ("main_beam", C.c_float), ("end_beam", C.c_float),
                ("word_beam", C.c_float), ("max_hyps", C.c_int32), ("device", C.c_int32),
                ("n_lanes", C.c_int32), ("max_active", C.c_int32), ("max_frames", C.c_int32),
                ("max_paths", C.c_int32), ("frame_stats", C.c_int32), ("reserved", C.c_int32)]


class JgpuWord(C.Structure):
    _fields_ = [("label", C.c_int32), ("time", C.c_int32), ("score", C.c_float), ("ac", C.c_float),
                ("lm", C.c_float)]


class JgpuResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("n_frames", C.c_int32), ("score", C.c_float), ("ac", C.c_float),
                ("lm", C.c_float), ("max_words", C.c_int32), ("words", C.POINTER(JgpuWord))]


class JgpuStats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_frames", "total_active_models", "total_active_emit_hyps",
                                         "total_active_end_hyps", "total_proc_emit_hyps", "total_proc_end_hyps",
                                         "total_gmm_evals", "total_arcs_expanded", "total_entry_writes",
                                         "total_paths")]

    def as_dict(self) -> Dict[str, int]:
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


class FlatTables:
    """Owns contiguous numpy arrays and the three C structs pointing into them."""

    def __init__(self, net: Dict[str, np.ndarray], init_state: int, models: Dict[str, np.ndarray]):
        c = np.ascontiguousarray
        self.arc_to = c(net["arc_to"], dtype=np.int32)
        self.arc_w = c(net["arc_w"], dtype=np.float32)
        self.arc_in = c(net["arc_in"], dtype=np.int32)
        self.arc_out = c(net["arc_out"], dtype=np.int32)
        self.st_first = c(net["st_first"], dtype=np.int32)
        self.st_n = c(net["st_n"], dtype=np.int32)
        self.st_final = c(net["st_final"], dtype=np.float32)
        self.hmm_nstates = c(models["hmm_nstates"], dtype=np.int32)
        self.hmm_gmm = c(models["hmm_gmm"], dtype=np.int32)
        self.trp = c(models["trP"], dtype=np.float32)
        self.se = c(models["se"], dtype=np.int32)
        self.tee = c(models["hmm_tee"], dtype=np.float32)
        self.gmm_ncomp = c(models["gmm_ncomp"], dtype=np.int32)
        self.dets = c(models["dets"], dtype=np.float32)
        self.means = c(models["means"], dtype=np.float32)
        self.ivars = c(models["ivars"], dtype=np.float32)
        self.net = JgpuNet(self.st_first.shape[0], self.arc_to.shape[0], int(init_state),
                           _p(self.arc_to, i32p), _p(self.arc_w, f32p), _p(self.arc_in, i32p),
                           _p(self.arc_out, i32p), _p(self.st_first, i32p), _p(self.st_n, i32p),
                           _p(self.st_final, f32p))
        H, S = self.hmm_gmm.shape
        self.hmm = JgpuHmm(H, S, _p(self.hmm_nstates, i32p), _p(self.hmm_gmm, i32p), _p(self.trp, f32p),
                           _p(self.se, i32p), _p(self.tee, f32p))
        G, Cc, D = self.means.shape
        self.gmm = JgpuGmm(G, D, Cc, _p(self.gmm_ncomp, i32p), _p(self.dets, f32p), _p(self.means, f32p),
                           _p(self.ivars, f32p))

    @property
    def dim(self) -> int:
        return int(self.means.shape[2])

    @property
    def n_gmm(self) -> int:
        return int(self.means.shape[0])


def make_cfg(*, main_beam: float, start_beam: float = 0.0, end_beam: float = 0.0, word_beam: float = 0.0,
             max_hyps: int = 0, device: int = 0, n_lanes: int = 1, max_active: int = 0, max_frames: int = 0,
             max_paths: int = 0, frame_stats: int = 0) -> JgpuCfg:
    return JgpuCfg(start_beam, main_beam, end_beam, word_beam, max_hyps, device, n_lanes, max_active,
                   max_frames, max_paths, frame_stats, 0)


def words_to_list(res: JgpuResult) -> List[Dict]:
    n = max(0, min(int(res.status), int(res.max_words)))
    return [dict(label=int(res.words[i].label), time=int(res.words[i].time),
                 score=float(res.words[i].score), ac=float(res.words[i].ac), lm=float(res.words[i].lm))
            for i in range(n)]
