"""Seeded synthetic fixtures for the decode path.

Writes exactly the on-disk formats the reference loads, so the SAME files feed the
reference oracle and the CUDA path:

* acoustic models  -> JMBI binary   (reference reader: src/HTKModels.cpp:1112-1245,
                                     record layouts :1309-1372, :1440-1492, :1580-1633,
                                     :1697-1739, :1805-1851, :2002-2089)
* network          -> AT&T FSM text + symbol tables (reference reader:
                                     src/WFSTNetwork.cpp:371-616, symbols :52-112)
* features         -> float32 [T, 39] sampled along a random accepted path, so every
                      utterance has a planted right answer.

numpy only; nothing here touches the GPU or the oracle.
"""
from __future__ import annotations

import dataclasses
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

D_FEAT = 39
LOG_2_PI = 1.83787706640934548355


# --------------------------------------------------------------------------------------
# acoustic models
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class TransMat:
    n_states: int
    sucs: List[List[int]]            # successors per state
    probs: List[List[float]]         # linear probabilities per successor


@dataclasses.dataclass
class Models:
    """HTK-style model set: GMM pool + HMM table + transition matrices."""
    dim: int
    weights: List[np.ndarray]        # per GMM: [M_g] float32 mixture weights
    means: List[np.ndarray]          # per GMM: [M_g, D] float32
    vars_: List[np.ndarray]          # per GMM: [M_g, D] float32 (variances, not inverse)
    tmats: List[TransMat]
    hmm_nstates: List[int]
    hmm_gmm: List[List[int]]         # per HMM: gmm id per state (-1 for entry / exit)
    hmm_tmat: List[int]
    hmm_names: List[str]

    @property
    def n_gmm(self) -> int:
        return len(self.weights)

    @property
    def n_hmm(self) -> int:
        return len(self.hmm_nstates)


def left_to_right_tmat(n_states: int = 5, self_loop: float = 0.6) -> TransMat:
    """Entry -> 1 with p=1; emitting states self-loop / forward; no tee."""
    sucs: List[List[int]] = [[1]]
    probs: List[List[float]] = [[1.0]]
    for s in range(1, n_states - 1):
        sucs.append([s, s + 1])
        probs.append([self_loop, 1.0 - self_loop])
    sucs.append([])
    probs.append([])
    return TransMat(n_states, sucs, probs)


def tee_tmat(p_skip: float = 0.3, self_loop: float = 0.6) -> TransMat:
    """3-state 'sp' model with an entry->exit tee transition.  The reference only sees a tee
    when the exit is successor index >= 1 of the entry state (src/HTKModels.cpp:1358-1370)."""
    return TransMat(3, [[1, 2], [1, 2], []], [[1.0 - p_skip, p_skip], [self_loop, 1.0 - self_loop], []])


def skip_tmat(n_states: int = 6) -> TransMat:
    """Left-to-right with skip transitions (wider SEIndex ranges, gaps filled with LOG_ZERO)."""
    sucs: List[List[int]] = [[1, 2]]
    probs: List[List[float]] = [[0.8, 0.2]]
    for s in range(1, n_states - 1):
        nxt = [s, s + 1] + ([s + 2] if s + 2 <= n_states - 1 else [])
        p = [0.5, 0.35, 0.15][: len(nxt)]
        tot = sum(p)
        sucs.append(nxt)
        probs.append([x / tot for x in p])
    sucs.append([])
    probs.append([])
    return TransMat(n_states, sucs, probs)


def make_models(n_hmm: int, n_mix: int, *, sigma_mu: float = 0.8, seed: int = 0,
                n_gmm_pool: Optional[int] = None, ragged_mix: bool = False,
                with_tee: bool = False, mixed_topology: bool = False) -> Models:
    """n_hmm physical HMMs (5 states, 3 emitting) over untied (gmm = 3*hmm + s) or tied GMMs.

    sigma_mu is the workload knob: smaller -> more confusable states -> more active
    hypotheses.  with_tee appends one 3-state tee model as the LAST
    HMM; mixed_topology makes every 7th HMM a 6-state skip model and every 5th a 4-state one
    (exercises variable nStates and SEIndex gaps).
    """
    rng = np.random.default_rng(seed)
    tmats = [left_to_right_tmat(5)]
    if mixed_topology:
        tmats += [skip_tmat(6), left_to_right_tmat(4)]
    hmm_nstates: List[int] = []
    hmm_tmat: List[int] = []
    for h in range(n_hmm):
        if mixed_topology and h % 7 == 3:
            hmm_nstates.append(6); hmm_tmat.append(1)
        elif mixed_topology and h % 5 == 2:
            hmm_nstates.append(4); hmm_tmat.append(2)
        else:
            hmm_nstates.append(5); hmm_tmat.append(0)
    if with_tee:
        tmats.append(tee_tmat())
        hmm_nstates.append(3); hmm_tmat.append(len(tmats) - 1)
    n_emit = sum(n - 2 for n in hmm_nstates)
    n_gmm = n_emit if n_gmm_pool is None else n_gmm_pool
    if n_gmm_pool is None:
        assign = np.arange(n_emit)
    else:
        assign = rng.integers(0, n_gmm, size=n_emit)
        assign[: min(n_gmm, n_emit)] = rng.permutation(n_gmm)[: min(n_gmm, n_emit)]
    hmm_gmm: List[List[int]] = []
    k = 0
    for n in hmm_nstates:
        row = [-1] + [int(assign[k + i]) for i in range(n - 2)] + [-1]
        k += n - 2
        hmm_gmm.append(row)
    weights, means, vars_ = [], [], []
    for g in range(n_gmm):
        m = n_mix if not ragged_mix else int(rng.integers(1, n_mix + 1))
        w = rng.dirichlet(np.full(m, 5.0)).astype(np.float32)
        weights.append(w)
        means.append((rng.standard_normal((m, D_FEAT)) * sigma_mu).astype(np.float32))
        vars_.append(rng.uniform(0.5, 2.0, size=(m, D_FEAT)).astype(np.float32))
    names = [f"h{h}" for h in range(len(hmm_nstates))]
    if with_tee:
        names[-1] = "sp"
    return Models(D_FEAT, weights, means, vars_, tmats, hmm_nstates, hmm_gmm, hmm_tmat, names)


def _rec(tag: bytes, *parts: bytes) -> bytes:
    return tag + b"".join(parts)


def _i32(*v: int) -> bytes:
    return np.asarray(v, dtype="<i4").tobytes()


def _name(s: Optional[str]) -> bytes:
    if not s:
        return _i32(0)
    b = s.encode() + b"\0"
    return _i32(len(b)) + b


def write_jmbi(m: Models, path: str) -> None:
    """One mean/var vector per Gaussian, one mixture per GMM in GMM order (HTKFlatModels
    assumes gMMs[i].mixtureInd == i, src/HTKFlatModels.cpp:143-176)."""
    D = m.dim
    n_gauss = int(sum(len(w) for w in m.weights))
    out: List[bytes] = [b"JMBI", _i32(D, n_gauss, n_gauss, m.n_gmm, m.n_gmm, len(m.tmats), m.n_hmm)]
    all_means = np.concatenate(m.means, axis=0).astype("<f4")
    all_vars = np.concatenate(m.vars_, axis=0).astype("<f4")
    # JMMN: tag, nameLen=0, means[D]
    rec = np.zeros(n_gauss, dtype=[("tag", "S4"), ("len", "<i4"), ("v", "<f4", (D,))])
    rec["tag"] = b"JMMN"; rec["v"] = all_means
    out.append(rec.tobytes())
    # JMVR: tag, nameLen=0, vars[D], minusHalfOverVars[D], gconst   (src/HTKModels.cpp:857-866)
    rec = np.zeros(n_gauss, dtype=[("tag", "S4"), ("len", "<i4"), ("v", "<f4", (D,)),
                                   ("mh", "<f4", (D,)), ("g", "<f4")])
    rec["tag"] = b"JMVR"; rec["v"] = all_vars
    rec["mh"] = (np.float32(-0.5) / all_vars).astype("<f4")
    g = np.full(n_gauss, np.float32(D * LOG_2_PI), dtype=np.float32)
    logv = np.log(all_vars.astype(np.float64))
    for d in range(D):                                  # same accumulation order, float each step
        g = (g.astype(np.float64) + logv[:, d]).astype(np.float32)
    rec["g"] = (g.astype(np.float64) * -0.5).astype(np.float32)
    out.append(rec.tobytes())
    # JMMX: tag, nameLen=0, nComps, meanInds[n], varInds[n]
    k = 0
    for w in m.weights:
        n = len(w)
        idx = np.arange(k, k + n, dtype="<i4")
        out.append(_rec(b"JMMX", _i32(0, n), idx.tobytes(), idx.tobytes()))
        k += n
    # JMGM: tag, nameLen=0, mixtureInd, nComps, w[n], logw[n]
    for gi, w in enumerate(m.weights):
        lw = np.log(w.astype(np.float64)).astype("<f4")
        out.append(_rec(b"JMGM", _i32(0, gi, len(w)), w.astype("<f4").tobytes(), lw.tobytes()))
    # JMTM: tag, nameLen=0, nStates, nSucs[N], sucs[total], probs[total], logProbs[total]
    for t in m.tmats:
        nsucs = [len(s) for s in t.sucs]
        sucs = [x for s in t.sucs for x in s]
        probs = np.asarray([x for p in t.probs for x in p], dtype="<f4")
        lp = np.log(probs.astype(np.float64)).astype("<f4")
        out.append(_rec(b"JMTM", _i32(0, t.n_states), _i32(*nsucs), _i32(*sucs), probs.tobytes(), lp.tobytes()))
    # JMHM: tag, name, nStates, gmmInds[N], transMatInd
    for h in range(m.n_hmm):
        out.append(_rec(b"JMHM", _name(m.hmm_names[h]), _i32(m.hmm_nstates[h]), _i32(*m.hmm_gmm[h]),
                        _i32(m.hmm_tmat[h])))
    out.append(b"\0")                                   # hybridMode = false
    with open(path, "wb") as f:
        f.write(b"".join(out))



def write_mmf(m: Models, path: str, *, upper: bool = True, var_floor_macro: bool = True, digits: int = 9) -> None:
    """The same model set as HTK MMF text, laid out the way HTK's HHEd writes it: one ~o header
    (<STREAMINFO> <VECSIZE> <NULLD> <MFCC_E_D_A> <DIAGC>), an (ignored) ~v variance-floor macro, ~t macros for
    transition matrices used by several HMMs, ~s macros for GMMs tied across several HMM states, then the ~h
    definitions; single-component states carry no <NUMMIXES>/<MIXTURE> lines; <GCONST> is written (the reader
    recomputes it, src/HTKModels.cpp:839-875).  `digits` significant digits: 9 round-trips float32 exactly,
    HTK itself prints 7 (%e)."""
    D = m.dim
    kw = (lambda s: f"<{s.upper()}>") if upper else (lambda s: f"<{s}>")
    fmt = f"%.{digits - 1}e"
    vec = lambda v: " " + " ".join(fmt % float(x) for x in v)          # noqa: E731
    gmm_uses = np.zeros(m.n_gmm, dtype=np.int64)
    for row in m.hmm_gmm:
        for g in row:
            if g >= 0:
                gmm_uses[g] += 1
    tm_uses = np.bincount(np.asarray(m.hmm_tmat), minlength=len(m.tmats))
    out: List[str] = []

    def emit_tmat(t: TransMat, indent: str) -> None:
        out.append(f"{indent}{kw('TransP')} {t.n_states}")
        dense = np.zeros((t.n_states, t.n_states), dtype=np.float32)
        for i, (ss, pp) in enumerate(zip(t.sucs, t.probs)):
            for j, p in zip(ss, pp):
                dense[i, j] = np.float32(p)
        for i in range(t.n_states):
            out.append(indent + vec(dense[i]))

    def emit_gmm(g: int, indent: str) -> None:
        w, mu, var = m.weights[g], m.means[g], m.vars_[g]
        n = len(w)
        if n > 1:
            out.append(f"{indent}{kw('NumMixes')} {n}")
        for c in range(n):
            if n > 1:
                out.append(f"{indent}{kw('Mixture')} {c + 1} " + fmt % float(w[c]))
            out.append(f"{indent}{kw('Mean')} {D}")
            out.append(indent + vec(mu[c]))
            out.append(f"{indent}{kw('Variance')} {D}")
            out.append(indent + vec(var[c]))
            gc = D * LOG_2_PI + float(np.sum(np.log(var[c].astype(np.float64))))
            out.append(f"{indent}{kw('GConst')} " + "%e" % gc)

    out.append("~o")
    out.append(f"{kw('StreamInfo')} 1 {D}")
    out.append(f"{kw('VecSize')} {D}{kw('NullD')}<MFCC_E_D_A>{kw('DiagC')}")
    if var_floor_macro:
        out.append('~v "varFloor1"')
        out.append(f"{kw('Variance')} {D}")
        out.append(vec(np.full(D, 0.01, dtype=np.float32)))
    for k, t in enumerate(m.tmats):
        if tm_uses[k] > 1:
            out.append(f'~t "T_{k}"')
            emit_tmat(t, "")
    for g in range(m.n_gmm):
        if gmm_uses[g] > 1:
            out.append(f'~s "ST_{g}"')
            emit_gmm(g, "")
    for h in range(m.n_hmm):
        out.append(f'~h "{m.hmm_names[h]}"')
        out.append(kw("BeginHMM"))
        out.append(f"{kw('NumStates')} {m.hmm_nstates[h]}")
        for s in range(1, m.hmm_nstates[h] - 1):
            g = m.hmm_gmm[h][s]
            out.append(f"{kw('State')} {s + 1}")
            if gmm_uses[g] > 1:
                out.append(f'~s "ST_{g}"')
            else:
                emit_gmm(g, "")
        k = m.hmm_tmat[h]
        if tm_uses[k] > 1:
            out.append(f'~t "T_{k}"')
        else:
            emit_tmat(m.tmats[k], "")
        out.append(kw("EndHMM"))
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")

# --------------------------------------------------------------------------------------
# networks
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Net:
    """Arcs in FILE order (grouped by source state, initial state's arcs first).
    w are FILE weights = -log probabilities (the loader negates, src/WFSTNetwork.cpp:481)."""
    src: np.ndarray
    dst: np.ndarray
    ilab: np.ndarray
    olab: np.ndarray
    w: np.ndarray
    finals: Dict[int, float]          # state -> file weight
    n_in: int                         # number of input symbols incl. <eps>
    n_out: int                        # number of output symbols incl. <eps>
    in_names: Optional[List[str]] = None
    out_names: Optional[List[str]] = None   # output symbol names by id (index 0 = <eps>); default W<id-1>

    @property
    def n_arcs(self) -> int:
        return int(self.src.shape[0])

    @property
    def n_states(self) -> int:
        return int(max(self.src.max(), self.dst.max())) + 1


def _finish_net(src, dst, ilab, olab, w, finals, n_in, n_out, in_names=None) -> Net:
    src = np.asarray(src, dtype=np.int64); dst = np.asarray(dst, dtype=np.int64)
    ilab = np.asarray(ilab, dtype=np.int64); olab = np.asarray(olab, dtype=np.int64)
    w = np.asarray(w, dtype=np.float64)
    order = np.argsort(src, kind="stable")             # group by source; state 0 (initial) first
    return Net(src[order], dst[order], ilab[order], olab[order], w[order], finals, n_in, n_out, in_names)


def digit_loop_net(n_words: int = 10) -> Net:
    """C1: one state, n self-loop arcs (in=h+1, out=h+1, w=ln n), final."""
    a = np.arange(n_words)
    z = np.zeros(n_words, dtype=np.int64)
    return _finish_net(z, z, a + 1, a + 1, np.full(n_words, math.log(n_words)), {0: 0.0},
                       n_words + 1, n_words + 1)


def tee_eps_net(n_words: int = 4, sp_label: Optional[int] = None) -> Net:
    """Second smoke fixture of SURVEY Appendix E: 0 -w:eps/ln n-> a_w -sp:eps-> b_w -eps:W/0.25-> 0,
    final 0 / 0.5.  Exercises tee skipping, eps arcs carrying word labels, final weights."""
    sp = sp_label if sp_label is not None else n_words + 1
    src, dst, il, ol, w = [], [], [], [], []
    for v in range(n_words):
        a, b = 1 + 2 * v, 2 + 2 * v
        src += [0, a, b]; dst += [a, b, 0]
        il += [v + 1, sp, 0]; ol += [0, 0, v + 1]
        w += [math.log(n_words), 0.0, 0.25]
    return _finish_net(src, dst, il, ol, w, {0: 0.5}, sp + 1, n_words + 1)


def ties_net() -> Net:
    """Exact score ties by construction (the tie-break test): a two-state loop whose weights are multiples of
    0.5, with homophones — words 0, 1 and 6 share HMM 0, words 2 and 3 share HMM 1 — and two routes of equal
    total weight to the same pronunciation: 0 -h0:W0/1-> 0 directly, or 0 -eps/0.5-> 1 -h0:W6/0.5-> 0.  Tokens on
    homophone arcs carry bit-identical scores, leave their models in the same frame and meet at state 0."""
    src = [0] * 7 + [1]
    dst = [0] * 6 + [1, 0]
    il = [1, 1, 2, 2, 3, 4, 0, 1]
    ol = [1, 2, 3, 4, 5, 6, 0, 7]
    w = [1.0] * 6 + [0.5, 0.5]
    return _finish_net(src, dst, il, ol, w, {0: 0.0}, 5, 8)


TIE_CLASSES = {1: 0, 2: 0, 7: 0, 3: 1, 4: 1, 5: 2, 6: 3}     # output label -> homophone class of ties_net()


def bigram_net(n_words: int, n_hmm: int, *, k_bigram: int = 8, seed: int = 0,
               pron_len: Tuple[int, int] = (3, 8), sp_label: Optional[int] = None,
               positive_weights: bool = True) -> Net:
    """C2-shaped C.L.G: unigram hub (state 0, initial+final) with one pronunciation chain per
    word (word label + LM weight on the first arc); per-word history states with k explicit
    bigram chains and one eps back-off arc to the hub.  With sp_label, every word chain ends
    in an optional-silence tee arc.  A few arcs get positive log-weights after negation
    (SURVEY 7 'Network weights are not guaranteed <= 0')."""
    rng = np.random.default_rng(seed)
    V = n_words
    prons = [rng.integers(1, n_hmm + 1, size=int(rng.integers(pron_len[0], pron_len[1] + 1))) for _ in range(V)]
    hist0 = 1                                            # history state of word v = 1 + v
    nxt = 1 + V
    src: List[int] = []; dst: List[int] = []; il: List[int] = []; ol: List[int] = []; w: List[float] = []

    def chain(s: int, v: int, e: int, lm: float) -> None:
        nonlocal nxt
        p = prons[v]
        cur = s
        n = len(p) + (1 if sp_label is not None else 0)
        for i in range(n):
            to = e if i == n - 1 else nxt
            if to == nxt:
                nxt += 1
            lab = int(p[i]) if i < len(p) else int(sp_label)
            src.append(cur); dst.append(to); il.append(lab)
            ol.append(v + 1 if i == 0 else 0); w.append(lm if i == 0 else 0.0)
            cur = to

    uni = -np.log(rng.dirichlet(np.full(V, 2.0)))
    for v in range(V):
        chain(0, v, hist0 + v, float(uni[v]))
    for v in range(V):
        succ = rng.choice(V, size=min(k_bigram, V), replace=False)
        pw = -np.log(rng.dirichlet(np.full(len(succ) + 1, 2.0)))
        for j, u in enumerate(succ):
            lm = float(pw[j])
            if positive_weights and rng.random() < 0.02:
                lm = -float(rng.uniform(0.05, 0.5))     # file weight < 0 -> positive log-weight
            chain(hist0 + v, int(u), hist0 + int(u), lm)
        src.append(hist0 + v); dst.append(0); il.append(0); ol.append(0); w.append(float(pw[-1]))
    finals = {0: 0.0}
    fw = rng.uniform(0.0, 1.0, size=V)
    for v in range(V):
        finals[hist0 + v] = float(fw[v])
    n_in = (n_hmm if sp_label is None else max(n_hmm, sp_label)) + 1
    return _finish_net(src, dst, il, ol, w, finals, n_in, V + 1)


def trigram_net(n_words: int, n_hmm: int, *, k_bigram: int = 40, n_trigram: int = 60000,
                k_trigram: int = 8, seed: int = 0, pron_len: Tuple[int, int] = (3, 8),
                prefix_tree: bool = False) -> Net:
    """C3/C5-shaped network, built vectorised:
      hub (state 0)        --first phone of w / out=w / unigram weight--> shared tail of w --> h_w
      bigram state h_v     --k_bigram first arcs into shared tails + eps back-off to hub
      trigram state t_(u,v): reached from h_u through a DEDICATED chain for v;
                             k_trigram first arcs into shared tails + eps back-off to h_v
    Shared tails make arcs >> states with a heavy-tailed out-degree (hub = V arcs); the
    eps back-off chain tri -> bi -> uni has depth 2.  All history states are final.
    prefix_tree=True replaces the hub's one-arc-per-word row by a PREFIX-TREE lexicon: words
    sharing their first phones share arcs, the hub's out-degree is the number of distinct first phones, the
    word label sits on the first arc that is unique to the word, and the unigram weights are pushed towards
    the root (an arc into a tree node carries min-cost(below it) - min-cost(below its parent)), the way a
    determinised and weight-pushed C.L.G looks.  After its unique arc a word runs into its shared tail."""
    rng = np.random.default_rng(seed)
    V = n_words
    plen = rng.integers(pron_len[0], pron_len[1] + 1, size=V)
    pron_off = np.concatenate([[0], np.cumsum(plen)])
    phones = rng.integers(1, n_hmm + 1, size=int(pron_off[-1]))
    # state numbering
    hist0 = 1                                        # h_v = 1 + v
    tail0 = hist0 + V                                # shared tail of w: states tail_base[w] .. + plen[w]-2
    tail_len = plen - 1
    tail_base = tail0 + np.concatenate([[0], np.cumsum(tail_len)])[:-1]
    tri0 = int(tail0 + tail_len.sum())               # t_k = tri0 + k
    tri_u = rng.integers(0, V, size=n_trigram)
    tri_v = rng.integers(0, V, size=n_trigram)
    ded_len = plen[tri_v] - 1                        # dedicated chain inner states for v from h_u
    ded0 = tri0 + n_trigram
    ded_base = ded0 + np.concatenate([[0], np.cumsum(ded_len)])[:-1]
    tree0 = int(ded0 + ded_len.sum())                # shared prefix-tree nodes (prefix_tree=True)

    S: List[np.ndarray] = []; T: List[np.ndarray] = []; I: List[np.ndarray] = []
    O: List[np.ndarray] = []; W: List[np.ndarray] = []

    def add(s, t, i, o, w):
        S.append(np.asarray(s, dtype=np.int64)); T.append(np.asarray(t, dtype=np.int64))
        I.append(np.asarray(i, dtype=np.int64)); O.append(np.asarray(o, dtype=np.int64))
        W.append(np.asarray(w, dtype=np.float64))

    def first_arcs(src_states: np.ndarray, words: np.ndarray, lm: np.ndarray):
        """arc src -> (tail_base[w] or h_w if 1-phone) labelled first phone of w, out = w+1."""
        to = np.where(tail_len[words] > 0, tail_base[words], hist0 + words)
        add(src_states, to, phones[pron_off[words]], words + 1, lm)

    # shared tails: for word w, phone i (1..plen-1): state tail_base+i-1 -> next (or h_w)
    w_rep = np.repeat(np.arange(V), tail_len)
    pos = np.arange(tail_len.sum()) - np.repeat(np.concatenate([[0], np.cumsum(tail_len)])[:-1], tail_len)
    s_ = tail_base[w_rep] + pos
    last = pos == tail_len[w_rep] - 1
    t_ = np.where(last, hist0 + w_rep, s_ + 1)
    add(s_, t_, phones[pron_off[w_rep] + pos + 1], np.zeros_like(s_), np.zeros(len(s_)))
    # hub
    uni = -np.log(rng.dirichlet(np.full(V, 2.0)))
    n_tree = 0
    if not prefix_tree:
        first_arcs(np.zeros(V, dtype=np.int64), np.arange(V), uni)
    else:
        # trie over the pronunciations; node ids are handed out after every other state (see tree0 below)
        kids: List[Dict[int, int]] = [{}]              # node -> {phone: child node}; node 0 = the hub
        cost = [math.inf]                              # min unigram cost of the words below the node
        cnt = [0]                                      # words below the node
        for w_ in range(V):
            node = 0
            cost[0] = min(cost[0], float(uni[w_])); cnt[0] += 1
            for i in range(int(plen[w_])):
                ph = int(phones[pron_off[w_] + i])
                nxt_ = kids[node].get(ph)
                if nxt_ is None:
                    nxt_ = len(kids)
                    kids[node][ph] = nxt_
                    kids.append({}); cost.append(math.inf); cnt.append(0)
                node = nxt_
                cost[node] = min(cost[node], float(uni[w_])); cnt[node] += 1
        cost[0] = 0.0                                  # nothing is pushed out of the hub
        t_src: List[int] = []; t_dst: List[int] = []; t_il: List[int] = []; t_ol: List[int] = []; t_w: List[float] = []
        tree_id: Dict[int, int] = {0: 0}               # trie node -> network state (shared nodes only)
        pending_nodes: List[int] = []
        for w_ in range(V):
            node = 0
            n_ph = int(plen[w_])
            for i in range(n_ph):
                ph = int(phones[pron_off[w_] + i])
                child = kids[node][ph]
                if cnt[child] > 1 and i < n_ph - 1:    # still shared: a tree arc, written once (by its first word)
                    if child not in tree_id:
                        tree_id[child] = -(len(pending_nodes) + 1)     # numbered below
                        pending_nodes.append(child)
                        t_src.append(tree_id[node]); t_dst.append(tree_id[child]); t_il.append(ph); t_ol.append(0)
                        t_w.append(cost[child] - cost[node])
                    node = child
                    continue
                # first arc unique to the word (or its last phone: homophones / prefixes of other words keep their
                # own arc there): word label + what is left of the unigram cost, then into the shared tail
                d_ = i + 1
                to = int(tail_base[w_] + d_ - 1) if d_ < n_ph else int(hist0 + w_)
                t_src.append(tree_id[node]); t_dst.append(to); t_il.append(ph); t_ol.append(w_ + 1)
                t_w.append(float(uni[w_]) - cost[node])
                break
        n_tree = len(pending_nodes)
    # bigram states
    kb = min(k_bigram, V)
    bw = rng.integers(0, V, size=(V, kb))
    lm = rng.gamma(4.0, 1.0, size=(V, kb))
    first_arcs(np.repeat(hist0 + np.arange(V), kb), bw.ravel(), lm.ravel())
    add(hist0 + np.arange(V), np.zeros(V), np.zeros(V), np.zeros(V), rng.gamma(2.0, 1.0, size=V))
    # dedicated chains h_u -> ... -> t_k spelling v (word label on the first arc)
    k_idx = np.arange(n_trigram)
    to0 = np.where(ded_len > 0, ded_base, tri0 + k_idx)
    add(hist0 + tri_u, to0, phones[pron_off[tri_v]], tri_v + 1, rng.gamma(3.0, 1.0, size=n_trigram))
    k_rep = np.repeat(k_idx, ded_len)
    pos = np.arange(ded_len.sum()) - np.repeat(np.concatenate([[0], np.cumsum(ded_len)])[:-1], ded_len)
    s_ = ded_base[k_rep] + pos
    last = pos == ded_len[k_rep] - 1
    t_ = np.where(last, tri0 + k_rep, s_ + 1)
    add(s_, t_, phones[pron_off[tri_v[k_rep]] + pos + 1], np.zeros_like(s_), np.zeros(len(s_)))
    # trigram states
    kt = min(k_trigram, V)
    tw = rng.integers(0, V, size=(n_trigram, kt))
    first_arcs(np.repeat(tri0 + k_idx, kt), tw.ravel(), rng.gamma(3.0, 1.0, size=n_trigram * kt))
    add(tri0 + k_idx, hist0 + tri_v, np.zeros(n_trigram), np.zeros(n_trigram), rng.gamma(2.0, 1.0, size=n_trigram))

    if prefix_tree:
        fix = lambda v: np.where(np.asarray(v) < 0, tree0 - 1 - np.asarray(v), np.asarray(v))   # noqa: E731  (-k-1 -> tree0 + k)
        add(fix(t_src), fix(t_dst), t_il, t_ol, t_w)

    finals: Dict[int, float] = {0: 0.0}
    fw = rng.uniform(0.0, 1.0, size=V + n_trigram)
    for v in range(V):
        finals[hist0 + v] = float(fw[v])
    for k in range(n_trigram):
        finals[tri0 + k] = float(fw[V + k])
    return _finish_net(np.concatenate(S), np.concatenate(T), np.concatenate(I), np.concatenate(O),
                       np.concatenate(W), finals, n_hmm + 1, V + 1)


def write_fsm(net: Net, prefix: str) -> Tuple[str, str, str]:
    """AT&T text + complete, dense symbol tables (SURVEY 8b preconditions)."""
    fsm, insyms, outsyms = prefix + ".fsm", prefix + ".insyms", prefix + ".outsyms"
    w = net.w.astype(np.float32)
    with open(fsm, "w") as f:
        # %.9g round-trips float32 exactly through sscanf("%f")
        lines = [f"{s} {t} {i} {o} {x:.9g}\n" for s, t, i, o, x in
                 zip(net.src.tolist(), net.dst.tolist(), net.ilab.tolist(), net.olab.tolist(), w.tolist())]
        f.write("".join(lines))
        for s, fw in net.finals.items():
            f.write(f"{s} {np.float32(fw):.9g}\n")
    with open(insyms, "w") as f:
        f.write("<eps> 0\n")
        for i in range(1, net.n_in):
            nm = net.in_names[i] if net.in_names else f"h{i - 1}"
            f.write(f"{nm} {i}\n")
    with open(outsyms, "w") as f:
        f.write("<eps> 0\n")
        for i in range(1, net.n_out):
            f.write(f"{net.out_names[i] if net.out_names else 'W%d' % (i - 1)} {i}\n")
    return fsm, insyms, outsyms


# --------------------------------------------------------------------------------------
# features with a planted answer
# --------------------------------------------------------------------------------------
class PathSampler:
    """Random accepted walks through a Net, emitting frames from the visited states' GMMs."""

    def __init__(self, net: Net, models: Models, tee_hmms: Sequence[int] = ()):
        self.net, self.m = net, models
        n = net.n_states
        self.first = np.zeros(n, dtype=np.int64)
        self.count = np.bincount(net.src, minlength=n)
        self.first[1:] = np.cumsum(self.count)[:-1]
        self.tee = set(int(h) for h in tee_hmms)
        self.init = int(net.src[0])

    def sample(self, min_frames: int, rng: np.random.Generator, dur: Tuple[int, int] = (2, 4),
               max_arcs: int = 100000) -> Tuple[np.ndarray, List[int]]:
        net, m = self.net, self.m
        state = self.init
        frames: List[np.ndarray] = []
        words: List[int] = []
        n_fr = 0
        for _ in range(max_arcs):
            if n_fr >= min_frames and state in net.finals:
                break
            c = int(self.count[state])
            if c == 0:
                raise RuntimeError("dead-end state in synthetic network")
            a = int(self.first[state] + rng.integers(0, c))
            il, ol = int(net.ilab[a]), int(net.olab[a])
            if ol:
                words.append(ol)
            if il:
                h = il - 1
                if not (h in self.tee and rng.random() < 0.5):
                    for s in range(1, m.hmm_nstates[h] - 1):
                        g = m.hmm_gmm[h][s]
                        d = int(rng.integers(dur[0], dur[1] + 1))
                        comp = rng.integers(0, len(m.weights[g]), size=d)
                        x = m.means[g][comp] + rng.standard_normal((d, m.dim)) * np.sqrt(m.vars_[g][comp])
                        frames.append(x.astype(np.float32))
                        n_fr += d
            state = int(net.dst[a])
        else:
            raise RuntimeError("walk did not terminate on a final state")
        if not frames:
            return np.zeros((0, m.dim), dtype=np.float32), words
        return np.ascontiguousarray(np.concatenate(frames, axis=0)), words


def make_fixture(name: str, outdir: str, models: Models, net: Net) -> Dict[str, str]:
    os.makedirs(outdir, exist_ok=True)
    jm = os.path.join(outdir, name + ".jmbi")
    write_jmbi(models, jm)
    fsm, ins, outs = write_fsm(net, os.path.join(outdir, name))
    return {"jmbi": jm, "fsm": fsm, "insyms": ins, "outsyms": outs}


# --------------------------------------------------------------------------------------
# named configurations (BASELINE.json configs + parity fixtures)
# --------------------------------------------------------------------------------------
def named_config(name: str):
    """Returns (models, net, tee_hmms, decoder_kwargs) for a named, seeded configuration.

    c1      BASELINE configs[0]: 10-word digit loop, 3-state monophones, 1-mix
    ties    homophones and equal-weight routes: exact score ties between different arrivals at one state
    tee     tee model + eps arcs with word labels + final weight
    mixed   4/5/6-state HMMs with skips, ragged mixtures, optional-silence tee, all four beams + histogram
    c2mini  small bigram network with tied states
    c2      BASELINE configs[1]: 1k-vocab bigram, 2000 triphone HMMs x 16-mix, ~42k states
    c3      BASELINE configs[2]: 20k-vocab trigram-shaped network, ~440k states / ~1.8M arcs, beam 250
    c3s     c3 topology at 1/8 scale (parity at sizes the CPU oracle finishes in seconds)
    c3p     c3 with a prefix-tree lexicon at the hub (shared first phones, pushed unigram weights); c3ps = 1/8 scale
    c5      BASELINE configs[4]: 64k-vocab trigram-shaped network, ~1.45M states / ~5.9M arcs (beam sweeps)
    """
    if name == "c1":
        return make_models(10, 1, sigma_mu=2.0, seed=1), digit_loop_net(10), (), dict(main_beam=200.0)
    if name == "c1h":
        return make_models(10, 1, sigma_mu=2.0, seed=1), digit_loop_net(10), (), dict(main_beam=200.0, max_hyps=12)
    if name == "nolabel":                       # final tokens without any word label -> inactive DecHyp
        net = digit_loop_net(4)
        net.olab[:] = 0
        return make_models(4, 1, sigma_mu=2.0, seed=5), net, (), dict(main_beam=200.0)
    if name == "ties":                          # exact ties between different arrivals (homophones, equal-weight routes)
        return make_models(4, 2, sigma_mu=1.5, seed=9), ties_net(), (), dict(main_beam=200.0)
    if name == "tee":
        return (make_models(4, 2, sigma_mu=2.0, seed=2, with_tee=True), tee_eps_net(4, sp_label=5), (4,),
                dict(main_beam=200.0))
    if name == "mixed":
        return (make_models(40, 3, sigma_mu=1.0, seed=7, with_tee=True, mixed_topology=True, ragged_mix=True),
                bigram_net(30, 40, k_bigram=4, seed=8, sp_label=41), (40,),
                dict(main_beam=150.0, end_beam=100.0, word_beam=80.0, start_beam=120.0, max_hyps=300))
    if name == "c2mini":
        return (make_models(120, 4, sigma_mu=0.9, seed=11, n_gmm_pool=150), bigram_net(100, 120, k_bigram=5, seed=12),
                (), dict(main_beam=180.0, end_beam=140.0))
    if name == "c2":
        return (make_models(2000, 16, sigma_mu=0.8, seed=3), bigram_net(1000, 2000, k_bigram=8, seed=4), (),
                dict(main_beam=200.0))
    if name == "c3":
        return (make_models(4000, 16, sigma_mu=1.3, seed=21, n_gmm_pool=6000),
                trigram_net(20000, 4000, k_bigram=40, n_trigram=60000, k_trigram=8, seed=22), (),
                dict(main_beam=250.0))
    if name == "c3p":                           # c3 with a prefix-tree lexicon at the hub
        return (make_models(4000, 16, sigma_mu=1.3, seed=21, n_gmm_pool=6000),
                trigram_net(20000, 4000, k_bigram=40, n_trigram=60000, k_trigram=8, seed=22, prefix_tree=True), (),
                dict(main_beam=250.0))
    if name == "c3ps":                          # the same at 1/8 scale (parity tests)
        return (make_models(600, 8, sigma_mu=1.0, seed=23, n_gmm_pool=900),
                trigram_net(2500, 600, k_bigram=20, n_trigram=6000, k_trigram=6, seed=24, prefix_tree=True), (),
                dict(main_beam=220.0))
    if name == "c5":                            # BASELINE configs[4]: 64k-vocab trigram, ~1.45M states / ~5.9M arcs;
        #                                             the beam and histogram settings are swept by the caller
        return (make_models(4000, 16, sigma_mu=1.3, seed=21, n_gmm_pool=6000),
                trigram_net(64000, 4000, k_bigram=40, n_trigram=200000, k_trigram=8, seed=52), (),
                dict(main_beam=250.0))
    if name == "c3s":
        return (make_models(600, 8, sigma_mu=1.0, seed=23, n_gmm_pool=900),
                trigram_net(2500, 600, k_bigram=20, n_trigram=6000, k_trigram=6, seed=24), (),
                dict(main_beam=220.0))
    raise KeyError(name)
