"""CPU suite (no GPU): pins the oracle port to the reference's golden vectors and, when the
compiled reference is present (oracle/_ref), to the reference itself on fresh random inputs;
checks the product's host loaders against the reference's in-memory tables."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, MMF_CASES, Golden, GoldenMmf, bits, flat_tables_from_files, flat_tables_from_mmf, same_result

from juicer_b200 import _abi, synth
from oracle import binding
from oracle.binding import OraclePort, OracleRef

HAVE_REF = os.path.exists(binding.REF_SO)


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_port_matches_golden(case, oracle_port_lib, product_lib):
    g = Golden(case)
    tabs, net, models = flat_tables_from_files(g.files)
    p = OraclePort(tabs, _abi.make_cfg(**g.kw))
    for u in range(g.n_utts):
        r = p.decode(g.feats(u), counters=True)
        g.check(u, r, r.frame_cnt, r.frame_best, "port")
    n = int(g.z["gmm_rows"])
    sc = p.gmm_scores(g.feats(0)[:n])
    assert np.array_equal(bits(sc), g.z["gmm"])
    p.close()


@pytest.mark.parametrize("remove_tee", [False, True])
@pytest.mark.parametrize("case", MMF_CASES)
def test_port_on_mmf_models_matches_golden(case, remove_tee, oracle_port_lib, product_lib):
    """Models read from MMF text by jgpu_load_mmf, decoded by the port: the reference's results on the same
    .mmf (HTKFlatModels::Load(mmf, removeInitialToFinalTransitions); expected_mmf.npz)."""
    g = GoldenMmf(case, remove_tee)
    tabs, net, models = flat_tables_from_mmf(g.files, remove_tee)
    p = OraclePort(tabs, _abi.make_cfg(**g.kw))
    for u in range(g.n_utts):
        r = p.decode(g.feats(u), counters=True)
        g.check(u, r, r.frame_cnt, r.frame_best, f"port on MMF models, remove_tee={remove_tee}")
    p.close()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("case", MMF_CASES)
def test_reference_on_mmf_models_reproduces_golden(case):
    g = GoldenMmf(case, False)
    o = OracleRef({k: v for k, v in g.files.items() if k != "jmbi"}, **g.kw)
    for u in range(g.n_utts):
        r = o.decode(g.feats(u), counters=True)
        g.check(u, r, r.frame_cnt, r.frame_best, "ref on MMF models")
    o.close()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_reference_reproduces_golden(case):
    g = Golden(case)
    o = OracleRef(g.files, **g.kw)
    for u in range(g.n_utts):
        r = o.decode(g.feats(u), counters=True)
        g.check(u, r, r.frame_cnt, r.frame_best, "ref")
    o.close()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("case,lm_scale,ins_pen", [("c1", 1.0, 0.0), ("tee", 1.3, -0.2), ("mixed", 0.8, 0.4),
                                                    ("c2mini", 1.0, -0.5)])
def test_host_loaders_match_reference_tables(case, lm_scale, ins_pen, product_lib):
    g = Golden(case)
    o = OracleRef(g.files, lm_scale=lm_scale, ins_penalty=ins_pen, **g.kw)
    tabs, net, models = flat_tables_from_files(g.files, lm_scale, ins_pen)
    rn, rm = o.dump_net(), o.dump_models()
    an, am = net.arrays(), models.arrays()
    assert net.init_state == o.init_state
    for k in rn:
        a, b = rn[k], an[k]
        if k == "st_first":
            m = rn["st_n"] > 0
            a, b = a[m], b[m]
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), k
    for k in am:
        assert np.array_equal(rm[k].view(np.uint32), am[k].view(np.uint32)), k
    o.close()


@pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/liboracle_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_port_matches_reference_random(seed, tmp_path, oracle_port_lib, product_lib):
    """Fresh random network/models/beams each seed: port == reference, bit for bit."""
    rng = np.random.default_rng(seed)
    n_hmm = int(rng.integers(12, 40))
    tee = bool(seed % 2)
    m = synth.make_models(n_hmm, int(rng.integers(1, 5)), sigma_mu=float(rng.uniform(0.7, 1.5)), seed=seed,
                          with_tee=tee, mixed_topology=bool(seed % 3 == 0), ragged_mix=True,
                          n_gmm_pool=None if seed == 1 else 3 * n_hmm // 2)
    net = synth.bigram_net(int(rng.integers(15, 50)), n_hmm, k_bigram=3, seed=seed + 50,
                           sp_label=(n_hmm + 1) if tee else None)
    files = synth.make_fixture("r", str(tmp_path), m, net)
    kw = dict(main_beam=float(rng.uniform(80, 200)), end_beam=float(rng.choice([0.0, 90.0])),
              word_beam=float(rng.choice([0.0, 70.0])), start_beam=float(rng.choice([0.0, 110.0])),
              max_hyps=int(rng.choice([0, 150])))
    o = OracleRef(files, **kw)
    tabs, _n, _m = flat_tables_from_files(files)
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    ps = synth.PathSampler(net, m, tee_hmms=[n_hmm] if tee else [])
    for u in range(3):
        x, _ = ps.sample(int(rng.integers(40, 160)), rng)
        a, b = o.decode(x, counters=True), p.decode(x, counters=True)
        same_result(a, b, f"seed{seed}/utt{u}")
        assert np.array_equal(a.frame_cnt[:, :5], b.frame_cnt[:, :5])
        assert np.array_equal(bits(a.frame_best), bits(b.frame_best))
    o.close(); p.close()


def test_planted_answer_recovered(tmp_path, oracle_port_lib, product_lib):
    """The generator plants a word sequence; the decoder must find it (SURVEY 8c ii)."""
    m, net, tee, kw = synth.named_config("c1")
    files = synth.make_fixture("c1", str(tmp_path), m, net)
    tabs, _n, _m = flat_tables_from_files(files)
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(3)
    for _ in range(3):
        x, words = ps.sample(90, rng)
        r = p.decode(x)
        assert r.labels == words
        # LM bookkeeping: every word costs ln(10) on the digit loop
        assert abs(r.lm + len(words) * np.log(10.0)) < 1e-3
    p.close()
