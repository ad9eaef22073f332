"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against
  (1) the committed golden vectors produced by the unmodified reference,
  (2) the oracle port on fresh seeded inputs at sizes the CPU finishes in seconds,
  (3) size-independent properties at the BASELINE configurations.
Integer/index work (labels, word-end frames, work counters) must be bit-exact.  Scores are
compared bit-for-bit as well; where noted, the fall-back tolerance is north_star's 1e-4."""
import numpy as np
import pytest

from helpers import GOLDEN_CASES, MMF_CASES, Golden, GoldenMmf, bits, flat_tables_from_files, flat_tables_from_mmf, same_result

from juicer_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu


def make_decoder(net, models, kw, **extra):
    return api.WFSTDecoderLite(net, models, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                               kw.get("word_beam", 0.0), kw.get("max_hyps", 0), **extra)


@pytest.fixture(scope="module")
def port_lib(oracle_port_lib, product_lib):
    return oracle_port_lib


# ---------------------------------------------------------------------------------------
# (1) golden vectors of the reference
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_gmm_scores_match_reference_golden(case, port_lib):
    """gmm_loglik KAT: every GMM x frame equals HTKFlatModels::calcOutput bit for bit."""
    g = Golden(case)
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw)
    n = int(g.z["gmm_rows"])
    sc = dec.gmm_scores(g.feats(0)[:n])
    ref = g.z["gmm"].view(np.float32)
    # double exp/log of CUDA and glibc are each <= 1 ulp but not identical: a last-bit difference after
    # narrowing is possible at ~2^-29 per call; none has been observed
    assert np.abs(sc - ref).max() <= 1e-5
    assert np.array_equal(bits(sc), g.z["gmm"])
    dec.close()


@pytest.mark.parametrize("lazy", [0, 1])
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_decode_matches_reference_golden(case, lazy, port_lib, monkeypatch):
    """lazy = 1: the opt-in per-step scorer of the stamped (GMM, lane) pairs (JUICER_B200_LAZY) instead of scoring
    every GMM 16 frames ahead; frame_stats arms its self-check."""
    monkeypatch.setenv("JUICER_B200_LAZY", str(lazy))
    g = Golden(case)
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=2, frame_stats=True)
    for u in range(g.n_utts):
        lane = u % 2
        r = dec.decode(g.feats(u), lane=lane)
        cnt, best = dec.frame_stats(lane)
        assert cnt.shape[0] == g.feats(u).shape[0]
        g.check(u, r, cnt, best, "gpu streaming")
    rb = dec.decode_batch([g.feats(u) for u in range(g.n_utts)])
    for u in range(g.n_utts):
        g.check(u, rb[u], what="gpu batch")
    dec.close()


@pytest.mark.parametrize("remove_tee", [False, True])
@pytest.mark.parametrize("case", MMF_CASES)
def test_decode_with_mmf_models_matches_reference_golden(case, remove_tee, port_lib):
    """Models read from HTK MMF text (jgpu_load_mmf) instead of JMBI: the CUDA path reproduces what the reference
    decodes after HTKFlatModels::Load(mmf, removeInitialToFinalTransitions) on the same file, bit for bit."""
    g = GoldenMmf(case, remove_tee)
    tabs, net, models = flat_tables_from_mmf(g.files, remove_tee)
    dec = make_decoder(net, models, g.kw, n_lanes=2, frame_stats=True)
    for u in range(g.n_utts):
        r = dec.decode(g.feats(u), lane=u % 2)
        cnt, best = dec.frame_stats(u % 2)
        g.check(u, r, cnt, best, f"gpu, MMF models, remove_tee={remove_tee}")
    dec.close()


# ---------------------------------------------------------------------------------------
# (2) oracle port on fresh inputs
# ---------------------------------------------------------------------------------------
@pytest.mark.parametrize("seed", [0, 1, 2, 3, 4, 5])
def test_random_networks_match_oracle(seed, tmp_path, port_lib):
    from oracle.binding import OraclePort
    rng = np.random.default_rng(100 + seed)
    n_hmm = int(rng.integers(12, 60))
    tee = bool(seed % 2)
    m = synth.make_models(n_hmm, int(rng.integers(1, 6)), sigma_mu=float(rng.uniform(0.6, 1.5)), seed=seed,
                          with_tee=tee, mixed_topology=bool(seed % 3 == 0), ragged_mix=bool(seed % 2 == 0),
                          n_gmm_pool=None if seed % 2 else 2 * n_hmm)
    if seed == 5:
        net = synth.trigram_net(120, n_hmm, k_bigram=6, n_trigram=150, k_trigram=3, seed=seed)
    else:
        net = synth.bigram_net(int(rng.integers(15, 80)), n_hmm, k_bigram=4, seed=seed + 50,
                               sp_label=(n_hmm + 1) if tee else None)
    files = synth.make_fixture("r", str(tmp_path), m, net)
    kw = dict(main_beam=float(rng.uniform(80, 220)), end_beam=float(rng.choice([0.0, 90.0])),
              word_beam=float(rng.choice([0.0, 70.0])), start_beam=float(rng.choice([0.0, 110.0])),
              max_hyps=int(rng.choice([0, 200])))
    tabs, netl, models = flat_tables_from_files(files)
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    dec = make_decoder(netl, models, kw, n_lanes=3, frame_stats=True)
    ps = synth.PathSampler(net, m, tee_hmms=[n_hmm] if tee else [])
    feats = [ps.sample(int(rng.integers(30, 200)), rng)[0] for _ in range(5)]
    sc = dec.gmm_scores(feats[0][:16])
    assert np.array_equal(bits(sc), bits(p.gmm_scores(feats[0][:16])))
    for u, x in enumerate(feats):
        a = p.decode(x, counters=True)
        b = dec.decode(x, lane=u % 3)
        same_result(a, b, f"seed{seed}/utt{u}")
        cnt, best = dec.frame_stats(u % 3)
        assert np.array_equal(a.frame_cnt[:, [0, 1, 2, 4]], cnt)
        assert np.array_equal(bits(a.frame_best), bits(best))
    for u, r in enumerate(dec.decode_batch(feats)):
        same_result(p.decode(feats[u]), r, f"seed{seed}/batch{u}")
    dec.close(); p.close()


def test_streaming_chunking_is_invisible(port_lib):
    """processFrame one frame at a time == all frames at once == batch (IDecoder contract)."""
    g = Golden("mixed")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=2)
    x = g.feats(1)
    whole = dec.decode(x)
    for chunk in (1, 7, 64):
        same_result(whole, dec.decode(x, lane=1, chunk=chunk), f"chunk={chunk}")
    g.check(1, whole)
    dec.close()


def test_interleaved_lanes_are_independent(port_lib):
    g = Golden("c2mini")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=2)
    a, b = g.feats(0), g.feats(1)
    dec.init(0); dec.init(1)
    ia = ib = 0
    while ia < a.shape[0] or ib < b.shape[0]:
        if ia < a.shape[0]:
            dec.process_frames(a[ia:ia + 5], 0); ia += 5
        if ib < b.shape[0]:
            dec.process_frames(b[ib:ib + 3], 1); ib += 3
    g.check(1, dec.finish(1))
    g.check(0, dec.finish(0))
    dec.close()


def test_ragged_batch_more_utterances_than_lanes(port_lib):
    """Lanes are refilled as utterances finish; empty and truncated utterances ride along."""
    from oracle.binding import OraclePort
    g = Golden("c2mini")
    tabs, net, models = flat_tables_from_files(g.files)
    p = OraclePort(tabs, _abi.make_cfg(**g.kw))
    m, snet, tee, kw = synth.named_config("c2mini")
    ps = synth.PathSampler(snet, m)
    rng = np.random.default_rng(9)
    feats = [ps.sample(int(rng.integers(20, 150)), rng)[0] for _ in range(11)]
    feats.insert(3, np.zeros((0, 39), np.float32))
    feats.insert(7, feats[0][:13])
    want = [p.decode(x) for x in feats]
    for lanes in (1, 4, 16):
        dec = make_decoder(net, models, g.kw, n_lanes=lanes)
        got = dec.decode_batch(feats)
        for u in range(len(feats)):
            same_result(want[u], got[u], f"lanes={lanes}/utt{u}")
        st = dec.stats(-1)
        assert st["n_frames"] == sum(x.shape[0] for x in feats)
        dec.close()
    p.close()


def test_device_resident_batch_equals_host_batch(port_lib):
    import torch
    g = Golden("mixed")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=2)
    feats = [g.feats(u) for u in range(g.n_utts)]
    host = dec.decode_batch(feats)
    n = np.asarray([f.shape[0] for f in feats], dtype=np.int32)
    off = np.concatenate([[0], np.cumsum(n)[:-1]]).astype(np.int64)
    packed = torch.from_numpy(np.concatenate(feats, axis=0)).cuda()
    torch.cuda.synchronize()
    devr = dec.decode_batch_device(packed.data_ptr(), off, n)
    for u in range(len(feats)):
        same_result(host[u], devr[u], f"utt{u}")
        g.check(u, devr[u])
    dec.close()


def test_run_to_run_determinism(port_lib):
    g = Golden("c2mini")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=1)
    first = dec.decode(g.feats(0))
    for _ in range(3):
        same_result(first, dec.decode(g.feats(0)))
    dec.close()


def test_capacity_overflow_fails_one_utterance_not_the_batch(port_lib):
    """A lane that runs out of instance slots reports a per-utterance failure; the next
    utterance on the same lane decodes correctly (failure handling)."""
    g = Golden("c2mini")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=1, max_active=64)
    r = dec.decode_batch([g.feats(0), g.feats(0)[:4]])
    assert r[0].status <= -10
    dec.close()
    dec = make_decoder(net, models, g.kw, n_lanes=1)
    g.check(0, dec.decode_batch([g.feats(0)])[0])
    dec.close()


def test_create_validates_what_the_reference_never_checks(port_lib):
    g = Golden("c1")
    tabs, net, models = flat_tables_from_files(g.files)
    tabs.arc_in[0] = 999                                  # not an HMM index + 1
    with pytest.raises(api.JuicerError, match="input label"):
        make_decoder(tabs, tabs, g.kw)
    tabs.arc_in[0] = 0
    tabs.arc_in[:] = 0                                    # epsilon self-loops: the reference would recurse forever
    with pytest.raises(api.JuicerError, match="cycle"):
        make_decoder(tabs, tabs, g.kw)


# ---------------------------------------------------------------------------------------
# (3) BASELINE-scale configurations: oracle on a bounded sample + size-independent properties
# ---------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c2_setup(tmp_path_factory, port_lib):
    m, net, tee, kw = synth.named_config("c2")
    d = tmp_path_factory.mktemp("c2")
    files = synth.make_fixture("c2", str(d), m, net)
    tabs, netl, models = flat_tables_from_files(files)
    return m, net, kw, tabs, netl, models


def test_c2_single_utterance_matches_oracle(c2_setup):
    """BASELINE configs[1]: 1k-vocab bigram, 16-mix triphones, ~42k states, one utterance."""
    from oracle.binding import OraclePort
    m, net, kw, tabs, netl, models = c2_setup
    ps = synth.PathSampler(net, m)
    x, words = ps.sample(250, np.random.default_rng(77))
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    want = p.decode(x, counters=True)
    dec = make_decoder(netl, models, kw, frame_stats=True)
    got = dec.decode(x)
    same_result(want, got, "c2")
    cnt, best = dec.frame_stats(0)
    assert np.array_equal(want.frame_cnt[:, [0, 1, 2, 4]], cnt)
    assert np.array_equal(bits(want.frame_best), bits(best))
    assert got.labels == words                             # planted answer
    sc = dec.gmm_scores(x[:32])
    assert np.array_equal(bits(sc), bits(p.gmm_scores(x[:32])))
    dec.close(); p.close()


def test_c2_histogram_and_beams_match_oracle(c2_setup):
    from oracle.binding import OraclePort
    m, net, kw, tabs, netl, models = c2_setup
    kw2 = dict(main_beam=200.0, max_hyps=6000, end_beam=150.0, word_beam=120.0, start_beam=170.0)
    ps = synth.PathSampler(net, m)
    x, _ = ps.sample(150, np.random.default_rng(78))
    p = OraclePort(tabs, _abi.make_cfg(**kw2))
    dec = make_decoder(netl, models, kw2, frame_stats=True)
    want, got = p.decode(x, counters=True), dec.decode(x)
    same_result(want, got, "c2 pruned")
    cnt, best = dec.frame_stats(0)
    assert np.array_equal(want.frame_cnt[:, [0, 1, 2, 4]], cnt)
    dec.close(); p.close()


def test_c2_batch_properties(c2_setup):
    """Size-independent properties on a batch the CPU could not decode in test time:
    planted answers are recovered, results do not depend on the lane count or on the
    position in the batch, and per-utterance decodes equal batch decodes."""
    m, net, kw, tabs, netl, models = c2_setup
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(79)
    pairs = [ps.sample(int(rng.integers(60, 220)), rng) for _ in range(40)]
    feats = [x for x, _ in pairs]
    d8 = make_decoder(netl, models, kw, n_lanes=8)
    r8 = d8.decode_batch(feats)
    assert sum(r.labels == w for r, (_, w) in zip(r8, pairs)) >= 38
    st = d8.stats(-1)
    assert st["n_frames"] == sum(x.shape[0] for x in feats)
    d3 = make_decoder(netl, models, kw, n_lanes=3)
    r3 = d3.decode_batch(feats[::-1])[::-1]
    for u in range(len(feats)):
        same_result(r8[u], r3[u], f"lanes 8 vs 3, utt{u}")
    one = d3.decode(feats[5])
    same_result(one, r8[5], "streaming vs batch")
    d8.close(); d3.close()


@pytest.mark.parametrize("name", ["c3s", "c3ps"])
def test_c3_scaled_trigram_matches_oracle(name, tmp_path, port_lib):
    """c3 topology (hub + shared tails + bigram/trigram back-off, hub out-degree = vocabulary)
    at 1/8 scale: exercises huge-state expansion and two-level epsilon back-off.  c3ps = the same with a
    prefix-tree lexicon at the hub (word labels and pushed unigram weights inside the tree)."""
    from oracle.binding import OraclePort
    m, net, tee, kw = synth.named_config(name)
    files = synth.make_fixture(name, str(tmp_path), m, net)
    tabs, netl, models = flat_tables_from_files(files)
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    dec = make_decoder(netl, models, kw, n_lanes=2, frame_stats=True)
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(80)
    for u in range(2):
        x, words = ps.sample(150, rng)
        want, got = p.decode(x, counters=True), dec.decode(x, lane=u)
        same_result(want, got, f"{name}/utt{u}")
        cnt, best = dec.frame_stats(u)
        assert np.array_equal(want.frame_cnt[:, [0, 1, 2, 4]], cnt)
        assert np.array_equal(bits(want.frame_best), bits(best))
    dec.close(); p.close()


def test_word_boundary_arena_is_garbage_collected(c2_setup):
    """collectPaths (src/WFSTDecoderLite.cpp:699-747): with an arena far smaller than the number of
    word-boundary records the utterances create, decodes still equal those of a large arena (which the
    tests above pin to the reference) — dead records are recycled by the mark + sweep collection — on
    the batch path, on a second pass over warm free lists, and on the streaming interface."""
    m, net, kw, tabs, netl, models = c2_setup
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(4242)
    feats = [ps.sample(int(n), rng)[0] for n in (260, 180, 90, 200)]
    big = make_decoder(netl, models, kw, n_lanes=2)
    want = big.decode_batch(feats)
    made = int(big.stats(-1)["total_paths"])
    big.close()
    small_arena = max(16384, made // 6)                     # the collection runs every 16 steps at half occupancy
    assert made > 3 * small_arena, f"the utterances must create many more records ({made}) than the arena holds"
    dec = make_decoder(netl, models, kw, n_lanes=2, max_paths=small_arena)
    for rep in range(2):                                     # second pass: arenas and free lists are re-used
        got = dec.decode_batch(feats)
        assert all(r.status > 0 for r in got), [r.status for r in got]
        for u in range(len(feats)):
            same_result(want[u], got[u], f"gc pass {rep} utt{u}")
        assert int(dec.stats(-1)["total_paths"]) == made
    same_result(want[0], dec.decode(feats[0], lane=1, chunk=37), "gc streaming")
    dec.close()


def test_lazy_scoring_evaluates_a_tight_superset(c2_setup, monkeypatch):
    """HTKFlatModels::calcOutput is only called for states a live token asks for (src/HTKFlatModels.cpp:226-262).
    With JUICER_B200_LAZY=1 the CUDA path scores, once per step, the (GMM, lane) pairs stamped one step ahead — a superset of what
    k_internal reads (its self-check, active with frame_stats, turns a missing stamp into a failed utterance) that
    must stay close to the exact set the oracle port counts, and well below scoring everything."""
    import ctypes as C
    from oracle.binding import OraclePort
    monkeypatch.setenv("JUICER_B200_LAZY", "1")
    m, net, kw, tabs, netl, models = c2_setup
    ps = synth.PathSampler(net, m)
    x, _ = ps.sample(200, np.random.default_rng(81))
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    p.lib.jor_debug_gmm_superset.restype = C.c_longlong
    p.lib.jor_debug_gmm_superset_reset()
    want = p.decode(x)
    exact = p.stats()["total_gmm_evals"]
    superset = int(p.lib.jor_debug_gmm_superset())
    dec = make_decoder(netl, models, kw, n_lanes=2, frame_stats=True)
    got = dec.decode(x, lane=1)
    same_result(want, got, "lazy scoring")
    evals = dec.stats(1)["total_gmm_evals"]
    dense = x.shape[0] * models.n_gmm
    assert exact <= evals <= dense
    assert evals <= 1.02 * superset + 64, (exact, superset, evals, dense)
    assert evals < 0.9 * dense, (exact, superset, evals, dense)
    dec.close(); p.close()


def test_streaming_partial_results_match_reference(tmp_path, port_lib):
    """jgpu_partial_result vs tracePartialPath (src/WFSTDecoderLite.cpp:822-897) of the unmodified reference, asked
    after every 10th frame of a c2 utterance: the converged words and their end frames are identical, and where the
    reference cannot trace (an instance without any word in its history) the CUDA path says so."""
    import os
    from oracle import binding
    if not os.path.exists(binding.REF_SO):
        pytest.skip("oracle/_ref did not travel to this box")
    m, net, tee, kw = synth.named_config("c2")
    files = synth.make_fixture("c2", str(tmp_path), m, net)
    tabs, netl, models = flat_tables_from_files(files)
    x, words = synth.PathSampler(net, m).sample(240, np.random.default_rng(90))
    ref = binding.OracleRef(files, **kw)
    want = ref.decode_partials(x, 10)
    assert sum(1 for _, lab, _ in want if lab) >= 5                  # the utterance does converge along the way
    dec = make_decoder(netl, models, kw, n_lanes=2)
    dec.init(1)
    assert dec.partial_result(1).status == -1                        # nothing decoded yet
    k = 0
    for t0 in range(0, x.shape[0], 10):
        dec.process_frames(x[t0:t0 + 10], lane=1)
        if t0 + 10 > x.shape[0]:
            break
        frame, labels, frames = want[k]
        assert frame == t0 + 9
        got = dec.partial_result(1)
        if labels is None:
            assert got.status == -3, (frame, got)
        else:
            assert got.status == len(labels) and got.labels == labels and got.times == frames, (frame, labels, frames, got)
        k += 1
    final = dec.finish(1)
    assert final.labels == words
    last = [w for w in want if w[1]][-1]
    assert final.labels[: len(last[1])] == last[1]                   # the partial results are a prefix of the final one
    dec.close(); ref.close()
