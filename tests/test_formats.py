"""On-disk formats next to the path: the JWNT binary network reader
(jgpu_load_jwnt <-> WFSTNetwork::readBinary, src/WFSTNetwork.cpp:1228-1370) against files written by the
reference's own WFSTNetwork::writeBinary (tests/golden/*/*.jwnt, tools/make_golden.py) and against the
reference's own reader (oracle/_ref); and the MMF text model reader (jgpu_load_mmf <-> HTKFlatModels::Load,
src/HTKModels.cpp:221-283, grammar src/htkparse.y.ypp, tokens src/htkparse.l.lpp) against the tables the reference
holds after loading the same files (tests/golden/*/expected_mmf.npz, tools/make_golden_mmf.py)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, MMF_CASES, Golden, GoldenMmf, same_model_tables

from juicer_b200 import api


def _same_tables(a, b, what):
    assert sorted(a) == sorted(b), what
    for k in a:
        x, y = np.asarray(a[k]), np.asarray(b[k])
        assert x.shape == y.shape, f"{what}: {k} shape"
        if x.dtype == np.float32:
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), f"{what}: {k}"
        else:
            assert np.array_equal(x, y), f"{what}: {k}"


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_jwnt_equals_text_network(case, product_lib):
    """A network written by the reference as JWNT and read back here is the text-loaded network, bit for bit
    (scale 1, no insertion penalty: writeBinary removes and readBinary re-applies both)."""
    g = Golden(case)
    text = api.WFSTNetwork(g.files["fsm"], g.files["insyms"], g.files["outsyms"])
    binary = api.WFSTNetwork.from_jwnt(os.path.join(g.dir, case + ".jwnt"))
    assert binary.init_state == text.init_state
    _same_tables(text.arrays(), binary.arrays(), case)


@pytest.mark.parametrize("case", ["mixed", "tee"])
def test_jwnt_reader_matches_reference_reader(case, product_lib):
    """Scaling factor and insertion penalty are applied after reading exactly like readBinary (:1351-1365)."""
    from oracle import binding
    binding.build(ref=True, port=False)
    if not os.path.exists(binding.REF_SO):
        pytest.skip("oracle/_ref is not built (no /root/reference here)")
    g = Golden(case)
    path = os.path.join(g.dir, case + ".jwnt")
    for scale, pen in [(1.0, 0.0), (1.3, -0.7), (0.5, 2.25)]:
        ref = binding.RefJwnt(path, scale, pen)
        mine = api.WFSTNetwork.from_jwnt(path, scale, pen)
        assert (ref.n_states, ref.n_arcs, ref.init_state) == (mine.c.n_states, mine.c.n_arcs, mine.init_state)
        _same_tables(ref.dump_net(), mine.arrays(), f"{case} scale={scale} pen={pen}")
        ref.close()


def test_jwnt_errors_are_loud(tmp_path, product_lib):
    g = Golden("tee")
    raw = open(os.path.join(g.dir, "tee.jwnt"), "rb").read()
    bad_id = tmp_path / "bad_id.jwnt"
    bad_id.write_bytes(b"XXXX" + raw[4:])
    with pytest.raises(api.JuicerError, match="invalid ID"):
        api.WFSTNetwork.from_jwnt(str(bad_id))
    cut = tmp_path / "cut.jwnt"
    cut.write_bytes(raw[: len(raw) // 2])
    with pytest.raises(api.JuicerError):
        api.WFSTNetwork.from_jwnt(str(cut))
    tail = tmp_path / "tail.jwnt"
    tail.write_bytes(raw[:-4] + b"JWNX")
    with pytest.raises(api.JuicerError, match=r"invalid ID \(2\)"):
        api.WFSTNetwork.from_jwnt(str(tail))
    with pytest.raises(api.JuicerError, match="opening"):
        api.WFSTNetwork.from_jwnt(str(tmp_path / "missing.jwnt"))


# ---------------------------------------------------------------------------------------
# MMF text models
# ---------------------------------------------------------------------------------------
def _ref_or_skip():
    from oracle import binding
    binding.build(ref=True, port=False)
    if not os.path.exists(binding.REF_SO):
        pytest.skip("oracle/_ref is not built (no /root/reference here)")
    return binding


@pytest.mark.parametrize("remove_tee", [False, True])
@pytest.mark.parametrize("case", MMF_CASES)
def test_mmf_tables_match_reference_golden(case, remove_tee, product_lib):
    """jgpu_load_mmf leaves the tables HTKFlatModels::Load(mmf, removeInitialToFinalTransitions) leaves, bit for
    bit: trP / SEIndex / tee weights, GMM order (shared ~s states first), recomputed gconst + log weights."""
    g = GoldenMmf(case, remove_tee)
    mine = api.HTKFlatModels.from_mmf(g.files["mmf"], remove_tee)
    same_model_tables(g.tables(), mine.arrays(), f"{case} remove_tee={remove_tee}")
    tee = mine.arrays()["hmm_tee"]
    assert (tee[-1] > np.float32(-1e30)) == (not remove_tee)        # the fixtures' last HMM is the tee model


def test_mmf_written_by_the_reference_is_read_back(product_lib):
    """tee.refout.mmf is HTKModels::output(f, false) of the JMBI fixture, i.e. text produced by the reference."""
    g = GoldenMmf("tee")
    mine = api.HTKFlatModels.from_mmf(os.path.join(g.dir, "tee.refout.mmf"))
    same_model_tables(g.tables("refout_tab_"), mine.arrays(), "tee.refout.mmf")


@pytest.mark.parametrize("case", MMF_CASES)
def test_mmf_parameters_equal_jmbi_parameters(case, product_lib):
    """The .mmf fixtures print 9 significant digits, so every parameter read from text is the float32 the JMBI
    file stores: topology tables are identical and, state by state, so are the Gaussians (the GMM numbering
    differs: text load puts shared states first)."""
    g = GoldenMmf(case)
    ma, mb = api.HTKFlatModels(g.files["jmbi"]), api.HTKFlatModels.from_mmf(g.files["mmf"])   # own the arrays' memory
    a, b = ma.arrays(), mb.arrays()
    for k in ("hmm_nstates", "trP", "se", "hmm_tee"):
        x, y = a[k], b[k]
        assert np.array_equal(x.view(np.uint32) if x.dtype == np.float32 else x,
                              y.view(np.uint32) if y.dtype == np.float32 else y), k
    for h in range(a["hmm_gmm"].shape[0]):
        for s in range(1, int(a["hmm_nstates"][h]) - 1):
            ga, gb = a["hmm_gmm"][h, s], b["hmm_gmm"][h, s]
            assert a["gmm_ncomp"][ga] == b["gmm_ncomp"][gb]
            n = a["gmm_ncomp"][ga]
            assert np.array_equal(a["means"][ga, :n].view(np.uint32), b["means"][gb, :n].view(np.uint32))
            assert np.array_equal(a["ivars"][ga, :n].view(np.uint32), b["ivars"][gb, :n].view(np.uint32))
            assert np.allclose(a["dets"][ga, :n], b["dets"][gb, :n], rtol=0, atol=2e-5)   # gconst: numpy log vs libm log


@pytest.mark.parametrize("cfg,style", [("c2mini", dict(upper=True)), ("c2mini", dict(upper=False, digits=7)),
                                       ("mixed", dict(upper=True, var_floor_macro=False)), ("c1", dict(upper=False))])
def test_mmf_reader_matches_reference_reader(cfg, style, tmp_path, product_lib):
    """Fresh model sets (tied states -> ~s macros, shared ~t matrices, both keyword spellings, HTK's 7-digit
    precision) through the reference's loader and through jgpu_load_mmf; also the reference's MMF -> JMBI
    conversion read back by jgpu_load_jmbi."""
    binding = _ref_or_skip()
    from juicer_b200 import synth
    m, _, _, _ = synth.named_config(cfg)
    mmf = str(tmp_path / "m.mmf")
    synth.write_mmf(m, mmf, **style)
    for remove_tee in (False, True):
        ref = binding.RefModels(mmf, remove_tee=remove_tee)
        mine = api.HTKFlatModels.from_mmf(mmf, remove_tee)
        same_model_tables(ref.dump_models(), mine.arrays(), f"{cfg} {style} remove_tee={remove_tee}")
        jm = str(tmp_path / f"conv{int(remove_tee)}.jmbi")
        ref.write(jm, True)
        conv = api.HTKFlatModels(jm)
        same_model_tables(ref.dump_models(), conv.arrays(), "reference MMF -> JMBI conversion")
        ref.close()


_MINI = """~o <VecSize> 2 <StreamInfo> 1 2 <MFCC> <DiagC> <NullD> <HmmSetId> set1
~t "T" <TransP> 3
 0 1 0
 0 .5 5e-1
 0 0 0
~s "S" <Mean> 2
 1 -2.5
 <Variance> 2
 1.0 +2
~h "a" <BeginHMM> <NumStates> 3 <State> 2 ~s "S" ~t "T" <EndHMM>
~h "b" <BeginHMM> <NumStates> 3 ~o <VecSize> 2 <State> 2 <NumMixes> 2
 <Mixture> 1 0.25 <Mean> 2 0 0 <Variance> 2 1 1 <GConst> 3.6
 <Mixture> 2 0.75 <Mean> 2 1 1 <Variance> 2 2 2
 <TransP> 3 0 1.0 0  0 0.9 0.1  0 0 0 <EndHMM>
"""


def test_mmf_token_and_grammar_corners(tmp_path, product_lib):
    """Integers inside real vectors, '.5' / '5e-1' / '+2' reals, macros and keywords on one line, ~o inside an HMM,
    a state without <NumMixes>, an omitted <GConst>: everything htkparse.l/.y accept."""
    p = tmp_path / "mini.mmf"
    p.write_text(_MINI)
    mini = api.HTKFlatModels.from_mmf(str(p))
    a = mini.arrays()
    assert a["hmm_nstates"].tolist() == [3, 3] and a["hmm_gmm"].tolist() == [[-1, 0, -1], [-1, 1, -1]]
    assert a["gmm_ncomp"].tolist() == [1, 2]
    assert a["means"][0, 0].tolist() == [1.0, -2.5] and a["ivars"][0, 0].tolist() == [1.0, 0.5]
    f32, f64 = np.float32, np.float64
    gc = f32(f64(f32(f64(f32(2 * 1.83787706640934548355)) + np.log(f64(1.0)))) + np.log(f64(2.0)))   # float accumulation
    assert a["dets"][0, 0] == f32(f64(gc) * -0.5)                                        # log weight of a 1-mix state: 0
    assert a["trP"][0, 1, 1] == f32(np.log(f64(f32(0.5)))) and a["trP"][0, 1, 2] == a["trP"][0, 1, 1]
    assert a["trP"][1, 1, 2] == f32(np.log(f64(f32(0.1))))
    assert a["se"][0].tolist() == [[0, 0], [0, 2], [1, 2]]
    try:
        binding = _ref_or_skip()
    except pytest.skip.Exception:
        return
    ref = binding.RefModels(str(p))
    same_model_tables(ref.dump_models(), a, "mini.mmf")
    ref.close()


@pytest.mark.parametrize("edit,msg", [
    (lambda t: t.replace('~s "S" <Mean>', '~u "S" <Mean>', 1), "syntax error"),                   # ~u macros are not in the grammar
    (lambda t: t.replace("<Mean> 2\n 1 -2.5", "<Mean> 3\n 1 -2.5 0", 1), "did not match global vec size"),
    (lambda t: t.replace(" 1 -2.5", " 1 -2.5 7", 1), "n_elems did not match"),
    (lambda t: t.replace("<Mixture> 1 0.25", "<Mixture> 1 1", 1), "weight must be written as a real"),
    (lambda t: t.replace('<State> 2 ~s "S"', '<State> 2 ~s "nope"', 1), "SMACRO string not found"),
    (lambda t: t.replace('~s "S" ~t "T"', '~s "S" ~t "U"', 1), "not found in htk_def"),
    (lambda t: t.replace('~h "a" <BeginHMM> <NumStates> 3', '~h "a" <BeginHMM> <NumStates> 4', 1), "did not match n_states"),
    (lambda t: t.replace("<TransP> 3 0 1.0 0  0 0.9 0.1  0 0 0", "<TransP> 3 0 1.0 0  0 0.9 0.1", 1), "TRANSP value"),
    (lambda t: t.replace("<VecSize> 2 <State>", "<VecSize> 3 <State>", 1), "does not equal NEW vec_size"),
    (lambda t: t.replace("<Mixture> 2 0.75 <Mean> 2 1 1 <Variance> 2 2 2\n", "", 1).replace("<NumMixes> 2", "<NumMixes> 1"),
     "compWeights[0] != 1.0"),
    (lambda t: "", "syntax error"),
])
def test_mmf_errors_are_loud(edit, msg, tmp_path, product_lib):
    """Every htkerror / error() of the reference's text path is a JGPU_E_IO with the message, not an exit()."""
    p = tmp_path / "bad.mmf"
    text = edit(_MINI)
    assert text != _MINI
    p.write_text(text)
    with pytest.raises(api.JuicerError, match=__import__("re").escape(msg)):
        api.HTKFlatModels.from_mmf(str(p))


def test_mmf_missing_file(product_lib, tmp_path):
    with pytest.raises(api.JuicerError, match="cannot open"):
        api.HTKFlatModels.from_mmf(str(tmp_path / "nope.mmf"))
