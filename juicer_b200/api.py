"""Python host binding of the C ABI (include/juicer_b200.h) — used by tests and bench.py.

The names follow the reference objects they stand in for:

* ``WFSTNetwork``      <- Juicer::WFSTNetwork(fsm, insyms, outsyms, lmScale, insPenalty)  (src/WFSTNetwork.h:112-124)
* ``HTKFlatModels``    <- Juicer::HTKFlatModels + readBinary                             (src/HTKFlatModels.h:29-40)
* ``WFSTDecoderLite``  <- Juicer::WFSTDecoderLite(network, models, phoneStartPruneWin, emitPruneWin,
                          phoneEndPruneWin, wordPruneWin, maxEmitHyps) with init / processFrame / finish
                          (src/WFSTDecoderLite.h:78-107), plus the batch entry points.

There is no CPU path: constructing a decoder without a CUDA device raises ``JuicerError``.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _abi
from ._abi import JgpuCfg, JgpuGmm, JgpuHmm, JgpuNet, JgpuResult, JgpuStats, JgpuWord

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjuicer_b200.so")
_lib = None


class JuicerError(RuntimeError):
    pass


def load_library() -> C.CDLL:
    """Loads the in-tree CUDA library; fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("JUICER_B200_LIB", LIB_PATH)      # instrumented build of the same sources (tools/)
    if not os.path.exists(path):
        raise JuicerError(f"{path} is missing: run `python -m juicer_b200.build` "
                          "(juicer_b200 has no fallback implementation)")
    lib = C.CDLL(path)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.jgpu_last_error.restype = C.c_char_p
    lib.jgpu_version.restype = C.c_char_p
    lib.jgpu_create.argtypes = [vp, vp, vp, vp, C.POINTER(vp)]
    lib.jgpu_destroy.argtypes = [vp]
    lib.jgpu_gmm_scores.argtypes = [vp, vp, i32, vp]
    lib.jgpu_utt_begin.argtypes = [vp, i32]
    lib.jgpu_push_frames.argtypes = [vp, i32, vp, i32]
    lib.jgpu_utt_end.argtypes = [vp, i32, vp]
    lib.jgpu_partial_result.argtypes = [vp, i32, vp]
    lib.jgpu_decode_batch.argtypes = [vp, vp, vp, i32, vp]
    lib.jgpu_decode_batch_device.argtypes = [vp, vp, vp, vp, i32, vp]
    lib.jgpu_decode_queue.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.jgpu_decode_queue_device.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.jgpu_ubench_fp32.argtypes = [vp, vp]
    lib.jgpu_retry_count.argtypes = [vp]
    lib.jgpu_retry_count.restype = i64
    lib.jgpu_stats.argtypes = [vp, i32, vp]
    lib.jgpu_frame_stats.argtypes = [vp, i32, vp, vp, i32]
    lib.jgpu_launch_count.argtypes = [vp]
    lib.jgpu_launch_count.restype = i64
    lib.jgpu_set_stream.argtypes = [vp, vp]
    lib.jgpu_profile.argtypes = [vp, i32]
    lib.jgpu_profile_read.argtypes = [vp, vp, vp]
    lib.jgpu_kernel_name.argtypes = [i32]
    lib.jgpu_kernel_name.restype = C.c_char_p
    lib.jgpu_load_fsm.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_float, C.c_float, vp]
    lib.jgpu_load_jwnt.argtypes = [C.c_char_p, C.c_float, C.c_float, vp]
    lib.jgpu_free_net.argtypes = [vp]
    lib.jgpu_load_jmbi.argtypes = [C.c_char_p, vp, vp]
    lib.jgpu_load_mmf.argtypes = [C.c_char_p, i32, vp, vp]
    lib.jgpu_free_models.argtypes = [vp, vp]
    _lib = lib
    return lib


def _check(rc: int, what: str) -> None:
    if rc < 0:
        raise JuicerError(f"{what} failed ({rc}): {load_library().jgpu_last_error().decode(errors='replace')}")


def _np(ptr, n: int, dtype) -> np.ndarray:
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class WFSTNetwork:
    """AT&T-text network loaded by the library's host loader (weights negated / scaled /
    insertion-penalised exactly like the reference loader)."""

    def __init__(self, fsm: str, insyms: str, outsyms: str, lm_scale: float = 1.0, ins_penalty: float = 0.0):
        self.lib = load_library()
        self.c = JgpuNet()
        _check(self.lib.jgpu_load_fsm(fsm.encode(), insyms.encode(), outsyms.encode(), lm_scale, ins_penalty,
                                      C.byref(self.c)), "jgpu_load_fsm")

    @classmethod
    def from_jwnt(cls, path: str, lm_scale: float = 1.0, ins_penalty: float = 0.0) -> "WFSTNetwork":
        """JWNT binary network (WFSTNetwork::readBinary, src/WFSTNetwork.cpp:1228-1370)."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.c = JgpuNet()
        _check(self.lib.jgpu_load_jwnt(path.encode(), lm_scale, ins_penalty, C.byref(self.c)), "jgpu_load_jwnt")
        return self

    def arrays(self) -> Dict[str, np.ndarray]:
        c = self.c
        A, S = c.n_arcs, c.n_states
        return dict(arc_to=_np(c.arc_to, A, np.int32), arc_w=_np(c.arc_weight, A, np.float32),
                    arc_in=_np(c.arc_in, A, np.int32), arc_out=_np(c.arc_out, A, np.int32),
                    st_first=_np(c.state_first, S, np.int32), st_n=_np(c.state_narcs, S, np.int32),
                    st_final=_np(c.state_final, S, np.float32))

    @property
    def init_state(self) -> int:
        return int(self.c.init_state)

    def __del__(self):
        try:
            self.lib.jgpu_free_net(C.byref(self.c))
        except Exception:
            pass


class HTKFlatModels:
    """JMBI model binary flattened like HTKFlatModels::init + createTrPandSEIndex."""

    def __init__(self, jmbi: str):
        self.lib = load_library()
        self.hmm, self.gmm = JgpuHmm(), JgpuGmm()
        _check(self.lib.jgpu_load_jmbi(jmbi.encode(), C.byref(self.hmm), C.byref(self.gmm)), "jgpu_load_jmbi")

    @classmethod
    def from_mmf(cls, mmf: str, remove_initial_to_final: bool = False) -> "HTKFlatModels":
        """HTK MMF text model definitions (HTKFlatModels::Load, src/HTKFlatModels.cpp:89-92)."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.hmm, self.gmm = JgpuHmm(), JgpuGmm()
        _check(self.lib.jgpu_load_mmf(mmf.encode(), int(remove_initial_to_final), C.byref(self.hmm), C.byref(self.gmm)),
               "jgpu_load_mmf")
        return self

    def arrays(self) -> Dict[str, np.ndarray]:
        h, g = self.hmm, self.gmm
        H, S, G, Cc, D = h.n_hmms, h.max_states, g.n_gmms, g.max_comps, g.dim
        return dict(hmm_nstates=_np(h.n_states, H, np.int32), hmm_gmm=_np(h.gmm, H * S, np.int32).reshape(H, S),
                    trP=_np(h.trp, H * S * S, np.float32).reshape(H, S, S),
                    se=_np(h.se, H * S * 2, np.int32).reshape(H, S, 2), hmm_tee=_np(h.tee, H, np.float32),
                    gmm_ncomp=_np(g.n_comps, G, np.int32), dets=_np(g.dets, G * Cc, np.float32).reshape(G, Cc),
                    means=_np(g.means, G * Cc * D, np.float32).reshape(G, Cc, D),
                    ivars=_np(g.ivars, G * Cc * D, np.float32).reshape(G, Cc, D))

    @property
    def dim(self) -> int:
        return int(self.gmm.dim)

    @property
    def n_gmm(self) -> int:
        return int(self.gmm.n_gmms)

    def __del__(self):
        try:
            self.lib.jgpu_free_models(C.byref(self.hmm), C.byref(self.gmm))
        except Exception:
            pass


class Result:
    """Mirror of DecHyp / DecHypHist chain (oldest word first)."""

    def __init__(self, res: JgpuResult):
        self.status = int(res.status)
        self.n_frames = int(res.n_frames)
        self.totals = np.asarray([res.score, res.ac, res.lm], dtype=np.float32)
        self.score, self.ac, self.lm = (float(x) for x in self.totals)
        self.words = _abi.words_to_list(res)

    @property
    def labels(self) -> List[int]:
        return [w["label"] for w in self.words]

    @property
    def times(self) -> List[int]:
        return [w["time"] for w in self.words]

    def __repr__(self) -> str:
        return (f"Result(status={self.status}, labels={self.labels}, times={self.times}, score={self.score!r}, "
                f"ac={self.ac!r}, lm={self.lm!r})")


class WFSTDecoderLite:
    """CUDA token-passing decoder behind the reference decoder's constructor signature."""

    def __init__(self, network, models, phone_start_beam: float = 0.0, main_beam: float = 0.0,
                 phone_end_beam: float = 0.0, word_beam: float = 0.0, max_hyps: int = 0, *, n_lanes: int = 1,
                 device: int = 0, max_active: int = 0, max_frames: int = 0, max_paths: int = 0,
                 frame_stats: bool = False, max_words: int = 256):
        self.lib = load_library()
        self.network, self.models = network, models       # keep the tables alive
        if isinstance(network, WFSTNetwork):
            net_c = network.c
        else:                                             # _abi.FlatTables
            net_c = network.net
        if isinstance(models, HTKFlatModels):
            hmm_c, gmm_c = models.hmm, models.gmm
        else:
            hmm_c, gmm_c = models.hmm, models.gmm
        self.dim = int(gmm_c.dim)
        self.n_gmm = int(gmm_c.n_gmms)
        self.n_lanes = n_lanes
        self.max_words = max_words
        self.cfg = _abi.make_cfg(main_beam=main_beam, start_beam=phone_start_beam, end_beam=phone_end_beam,
                                 word_beam=word_beam, max_hyps=max_hyps, device=device, n_lanes=n_lanes,
                                 max_active=max_active, max_frames=max_frames, max_paths=max_paths,
                                 frame_stats=int(frame_stats))
        self.h = C.c_void_p()
        _check(self.lib.jgpu_create(C.byref(net_c), C.byref(hmm_c), C.byref(gmm_c), C.byref(self.cfg),
                                    C.byref(self.h)), "jgpu_create")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.jgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- IModels::calcOutput over all GMMs ----
    def gmm_scores(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.zeros((x.shape[0], self.n_gmm), dtype=np.float32)
        _check(self.lib.jgpu_gmm_scores(self.h, x.ctypes.data, x.shape[0], out.ctypes.data), "jgpu_gmm_scores")
        return out

    # ---- IDecoder: init / processFrame / finish ----
    def init(self, lane: int = 0) -> None:
        _check(self.lib.jgpu_utt_begin(self.h, lane), "jgpu_utt_begin")

    def process_frames(self, x: np.ndarray, lane: int = 0) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.dim)
        _check(self.lib.jgpu_push_frames(self.h, lane, x.ctypes.data, x.shape[0]), "jgpu_push_frames")

    def partial_result(self, lane: int = 0) -> Result:
        """The words every live hypothesis of the running utterance agrees on (jgpu_partial_result)."""
        words = (JgpuWord * self.max_words)()
        res = JgpuResult(0, 0, 0.0, 0.0, 0.0, self.max_words, C.cast(words, C.POINTER(JgpuWord)))
        _check(self.lib.jgpu_partial_result(self.h, lane, C.byref(res)), "jgpu_partial_result")
        return Result(res)

    def finish(self, lane: int = 0) -> Result:
        words = (JgpuWord * self.max_words)()
        res = JgpuResult(0, 0, 0.0, 0.0, 0.0, self.max_words, C.cast(words, C.POINTER(JgpuWord)))
        _check(self.lib.jgpu_utt_end(self.h, lane, C.byref(res)), "jgpu_utt_end")
        return Result(res)

    def decode(self, x: np.ndarray, lane: int = 0, chunk: int = 0) -> Result:
        self.init(lane)
        if chunk <= 0:
            self.process_frames(x, lane)
        else:
            for i in range(0, x.shape[0], chunk):
                self.process_frames(x[i:i + chunk], lane)
        return self.finish(lane)

    # ---- batch ----
    def _result_buffers(self, n: int):
        words = (JgpuWord * (self.max_words * max(n, 1)))()
        res = (JgpuResult * max(n, 1))()
        base = C.addressof(words)
        for u in range(n):
            res[u].max_words = self.max_words
            res[u].words = C.cast(base + u * self.max_words * C.sizeof(JgpuWord), C.POINTER(JgpuWord))
        return words, res

    def decode_batch(self, feats: Sequence[np.ndarray]) -> List[Result]:
        feats = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1, self.dim) for f in feats]
        n = len(feats)
        ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in feats])
        nfr = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
        words, res = self._result_buffers(n)
        _check(self.lib.jgpu_decode_batch(self.h, ptrs, nfr.ctypes.data, n, res), "jgpu_decode_batch")
        return [Result(res[u]) for u in range(n)]

    def decode_batch_device(self, d_feats_ptr: int, row_offset: np.ndarray, n_frames: np.ndarray,
                            want_results: bool = True) -> List[Result]:
        row_offset = np.ascontiguousarray(row_offset, dtype=np.int64)
        n_frames = np.ascontiguousarray(n_frames, dtype=np.int32)
        n = int(n_frames.shape[0])
        words, res = self._result_buffers(n)
        _check(self.lib.jgpu_decode_batch_device(self.h, C.c_void_p(d_feats_ptr), row_offset.ctypes.data,
                                                 n_frames.ctypes.data, n, res), "jgpu_decode_batch_device")
        return [Result(res[u]) for u in range(n)] if want_results else []

    # ---- whole-utterance work stealing (BASELINE configs[3]) ----
    def decode_queue(self, queue, feats: Sequence[np.ndarray], order: Optional[np.ndarray] = None):
        """Decodes what this rank can claim of `feats` from the shared `queue` (dist.SharedQueue), utterance by
        utterance as lanes run dry.  Returns ({utterance index: Result}, busy_ms)."""
        feats = [np.ascontiguousarray(f, dtype=np.float32).reshape(-1, self.dim) for f in feats]
        n = len(feats)
        ptrs = (C.c_void_p * max(n, 1))(*[f.ctypes.data for f in feats])
        nfr = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
        return self._decode_queue(queue, n, nfr, order,
                                  lambda res, cl, ncl, busy, op: self.lib.jgpu_decode_queue(
                                      self.h, queue.q, ptrs, nfr.ctypes.data, n, op, res, cl, ncl, busy), "jgpu_decode_queue")

    def decode_queue_device(self, queue, d_feats_ptr: int, row_offset: np.ndarray, n_frames: np.ndarray,
                            order: Optional[np.ndarray] = None):
        row_offset = np.ascontiguousarray(row_offset, dtype=np.int64)
        nfr = np.ascontiguousarray(n_frames, dtype=np.int32)
        n = int(nfr.shape[0])
        return self._decode_queue(queue, n, nfr, order,
                                  lambda res, cl, ncl, busy, op: self.lib.jgpu_decode_queue_device(
                                      self.h, queue.q, C.c_void_p(d_feats_ptr), row_offset.ctypes.data, nfr.ctypes.data, n, op,
                                      res, cl, ncl, busy), "jgpu_decode_queue_device")

    def _decode_queue(self, queue, n, nfr, order, call, what):
        words, res = self._result_buffers(n)
        claimed = np.zeros(max(n, 1), dtype=np.int32)
        n_claimed = C.c_int32(0)
        busy = C.c_double(0.0)
        op = None
        if order is not None:
            order = np.ascontiguousarray(order, dtype=np.int32)
            op = order.ctypes.data
        _check(call(res, claimed.ctypes.data, C.byref(n_claimed), C.byref(busy), op), what)
        return {int(u): Result(res[int(u)]) for u in claimed[: n_claimed.value]}, float(busy.value)

    def fp32_peak_tops(self) -> float:
        """Non-FMA FP32 issue peak of the device, 10^12 op/s (jgpu_ubench_fp32)."""
        v = C.c_double(0.0)
        _check(self.lib.jgpu_ubench_fp32(self.h, C.byref(v)), "jgpu_ubench_fp32")
        return float(v.value)

    @property
    def retry_count(self) -> int:
        """Second passes (larger per-lane arenas, fewer lanes) run by the batch entry points so far."""
        return int(self.lib.jgpu_retry_count(self.h))

    # ---- counters ----
    def stats(self, lane: int = -1) -> Dict[str, int]:
        s = JgpuStats()
        _check(self.lib.jgpu_stats(self.h, lane, C.byref(s)), "jgpu_stats")
        return s.as_dict()

    def frame_stats(self, lane: int = 0, max_frames: int = 4096):
        cnt = np.zeros((max_frames, 4), dtype=np.int32)
        best = np.zeros(max_frames, dtype=np.float32)
        n = self.lib.jgpu_frame_stats(self.h, lane, cnt.ctypes.data, best.ctypes.data, max_frames)
        _check(n, "jgpu_frame_stats")
        return cnt[:n], best[:n]

    def set_stream(self, cuda_stream: int) -> None:
        _check(self.lib.jgpu_set_stream(self.h, C.c_void_p(cuda_stream)), "jgpu_set_stream")

    def profile(self, enable: bool) -> None:
        _check(self.lib.jgpu_profile(self.h, int(enable)), "jgpu_profile")

    def profile_read(self) -> Dict[str, Dict[str, float]]:
        ms = (C.c_double * 16)()
        cnt = (C.c_int64 * 16)()
        n = self.lib.jgpu_profile_read(self.h, ms, cnt)
        _check(n, "jgpu_profile_read")
        return {self.lib.jgpu_kernel_name(k).decode(): dict(ms=float(ms[k]), launches=int(cnt[k])) for k in range(n)}

    @property
    def launch_count(self) -> int:
        return int(self.lib.jgpu_launch_count(self.h))
