"""The callers either side of the decode path (juicer_b200/harness.py): HTK feature files, extended file
names, result-word extraction and the output formats of the reference's batch harness.  First the pieces against
strings derived by hand from the cited lines of src/DecoderSingleTest.cpp and src/DecoderBatchTest.cpp; then the
whole thing against the reference's OWN harness: DecoderBatchTest / DecoderSingleTest / DecVocabulary compiled
unmodified (oracle/harness_driver.cpp, `make -C oracle harness`) behind stand-ins for the two out-of-tree
libraries they use (Tracter's HTKSource / FrameSink, Torch3's DiskXFile / EditDistance)."""
import os
import struct

import numpy as np
import pytest

from juicer_b200 import harness as H


def test_extended_filenames():
    e = H.parse_extended_filename("utt1.mfc")
    assert (e.test_name, e.data_file, e.start, e.end) == ("utt1.mfc", "utt1.mfc", -1, -1)
    e = H.parse_extended_filename("spk_a_001=/data/long.mfc[120,480]")
    assert (e.test_name, e.data_file, e.start, e.end) == ("spk_a_001", "/data/long.mfc", 120, 480)
    e = H.parse_extended_filename("name=/data/f.mfc")
    assert (e.test_name, e.data_file, e.start, e.end) == ("name", "/data/f.mfc", -1, -1)
    for bad, msg in [("n=f[5,5]", "extStartFrame >= extEndFrame"), ("n=f[-1,4]", "extStartFrame < 0"),
                     ("n=f[0,0]", "extEndFrame <= 0"), ("n=f[3", "end frame"), ("n=f[x,4]", "start frame"),
                     ("n=", "real filename")]:
        with pytest.raises(H.HarnessError, match=msg):
            H.parse_extended_filename(bad)


def test_htk_round_trip_and_header(tmp_path):
    x = np.random.default_rng(0).normal(size=(57, 39)).astype(np.float32)
    p = str(tmp_path / "a.mfc")
    H.write_htk(p, x, samp_period=100000)
    raw = open(p, "rb").read()
    assert struct.unpack(">iihh", raw[:12]) == (57, 100000, 156, 6 | 0o400 | 0o1000)
    assert raw[12:16] == struct.pack(">f", float(x[0, 0]))            # big-endian float32 frames
    y, period, kind = H.read_htk(p)
    assert period == 100000 and kind == (6 | 0o400 | 0o1000)
    assert y.dtype == np.float32 and np.array_equal(y.view(np.uint32), x.view(np.uint32))
    open(str(tmp_path / "short.mfc"), "wb").write(raw[:200])
    with pytest.raises(H.HarnessError, match="expected"):
        H.read_htk(str(tmp_path / "short.mfc"))
    open(str(tmp_path / "c.mfc"), "wb").write(struct.pack(">iihh", 1, 100000, 4, 6 | 0o2000) + b"\0" * 4)
    with pytest.raises(H.HarnessError, match="compressed"):
        H.read_htk(str(tmp_path / "c.mfc"))


def test_frames_fed_to_the_decoder():
    # whole file
    assert H.frames_decoded(300) == (0, 300)
    assert H.frames_decoded(7) == (0, 7)
    assert H.frames_decoded(0) == (0, 0)
    # segment: end frame inclusive
    assert H.frames_decoded(300, 100, 199) == (100, 100)
    assert H.frames_decoded(300, 100, 1000) == (100, 200)             # clipped by the file
    # the 20-frame pre-read ignores the end frame (src/DecoderSingleTest.cpp:272-275)
    assert H.frames_decoded(300, 10, 14) == (10, 20)
    assert H.frames_decoded(25, 10, 14) == (10, 15)


def test_load_utterance_uses_the_segment(tmp_path):
    x = np.arange(200 * 3, dtype=np.float32).reshape(200, 3)
    p = str(tmp_path / "f.mfc")
    H.write_htk(p, x)
    ext, y, period = H.load_utterance(f"seg7={p}[50,149]", expected_dim=3)
    assert ext.test_name == "seg7" and np.array_equal(y, x[50:150]) and period == 100000
    with pytest.raises(H.HarnessError, match="vector size"):
        H.load_utterance(p, expected_dim=39)


def test_result_words_are_float32_differences():
    f = np.float32
    labels, times = [3, 1, 4, 2], [9, 30, 31, 77]
    ac = [f(-101.25), f(-350.5), f(-377.125), f(-901.0625)]
    lm = [f(-2.5), f(-7.25), f(-7.75), f(-12.0)]
    r = H.extract_result_words(labels, times, ac, lm)
    assert [w.index for w in r] == [2, 0, 3, 1]
    assert [(w.start_time, w.end_time) for w in r] == [(0, 9), (9, 30), (30, 31), (31, 77)]
    assert [w.acoustic_score for w in r] == [ac[0], f(ac[1] - ac[0]), f(ac[2] - ac[1]), f(ac[3] - ac[2])]
    assert [w.lm_score for w in r] == [lm[0], f(lm[1] - lm[0]), f(lm[2] - lm[1]), f(lm[3] - lm[2])]
    # sentence marks removed: the differences skip over them (src/DecoderSingleTest.cpp:441-462)
    r = H.extract_result_words(labels, times, ac, lm, sent_start_index=2, sent_end_index=1, remove_sent_marks=True)
    assert [w.index for w in r] == [0, 3]
    assert [(w.start_time, w.end_time) for w in r] == [(0, 30), (30, 31)]
    assert r[0].acoustic_score == ac[1] and r[1].acoustic_score == f(ac[2] - ac[1])
    assert H.extract_result_words([], [], [], []) == []


WORDS = ["one", "two", "three", "four"]


def _result():
    f = np.float32
    return H.extract_result_words([3, 1, 4], [9, 30, 76], [f(-100.5), f(-350.25), f(-900.75)],
                                  [f(-2.0), f(-4.5), f(-7.0)])


def test_output_formats():
    r = _result()
    assert H.format_result("ref", "/d/utt_01.mfc", WORDS, r, 77) == "three one four \n"
    assert H.format_result("trans", "/d/utt_01.mfc", WORDS, r, 77) == "three one four (trans-3)\n"
    assert H.format_result("mlf", "/d/utt_01.mfc", WORDS, r, 77) == '"*/utt_01.rec"\nthree\none\nfour\n.\n'
    # xmlf: 100 ns units; non-zero times get one frame added (src/DecoderBatchTest.cpp:377-398);
    # score = acoustic + lm of the word, %f
    want = ('"*/utt_01.rec"\n'
            "0 1000000 three -102.500000\n"
            "1000000 3100000 one -252.250000\n"
            "3100000 7700000 four -553.000000\n"
            ".\n")
    assert H.format_result("xmlf", "/d/utt_01.mfc", WORDS, r, 77) == want
    assert (H.format_result("verbose", "sym", WORDS, r, 77, expected=[2, -1, 3]) ==
            "sym\n\tExpected :  three <OOV> four \n\tActual :    three one four   [ 10 31 77 (77) ]\n")
    assert H.format_result("verbose", "sym", WORDS, r, 77) == "sym\n\tActual :    three one four   [ 10 31 77 (77) ]\n"
    # no result: what the nResultWords == 0 branch prints
    assert H.format_result("ref", "x", WORDS, [], 5) == "\n"
    assert H.format_result("mlf", "a/b.c.mfc", WORDS, [], 5) == '"*/b.c.rec"\n.\n'
    assert H.mlf_header() == "#!MLF!#\n"
    with pytest.raises(H.HarnessError):
        H.format_result("json", "x", WORDS, r, 1)


@pytest.mark.gpu
def test_decode_files_matches_decode_batch(tmp_path, product_lib):
    """HTK files + extended names through the harness == the same frames through decode_batch."""
    from helpers import Golden, flat_tables_from_files
    from test_gpu_parity import make_decoder
    g = Golden("mixed")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=2)
    feats = [g.feats(u) for u in range(2)]
    long = np.concatenate([feats[0], feats[1]], axis=0)
    p0, pl = str(tmp_path / "u0.mfc"), str(tmp_path / "long.mfc")
    H.write_htk(p0, feats[0])
    H.write_htk(pl, long)
    n0 = feats[0].shape[0]
    specs = [p0, f"second={pl}[{n0},{long.shape[0] - 1}]"]
    n_out = int(np.asarray(net.arrays()["arc_out"]).max())
    words = [f"w{i}" for i in range(n_out)]
    text = H.decode_files(dec, specs, words, fmt="xmlf", expected_dim=feats[0].shape[1])
    want = H.mlf_header()
    for name, x, r in zip(["u0.mfc", "second"], feats, dec.decode_batch(feats)):
        g_words = H.extract_result_words(r.labels, r.times, [w["ac"] for w in r.words], [w["lm"] for w in r.words])
        want += H.format_result("xmlf", name, words, g_words, x.shape[0])
    assert text == want and text.count(".rec") == 2
    for u in range(2):
        g.check(u, dec.decode_batch(feats)[u], what="harness")
    dec.close()


# ---------------------------------------------------------------------------------------
# pinned against the reference's own harness (oracle/_ref/liboracle_harness.so: DecoderBatchTest / DecoderSingleTest /
# DecVocabulary compiled unmodified, `make -C oracle harness`)
# ---------------------------------------------------------------------------------------
def _harness_fixture(tmp_path):
    """Small bigram network whose output symbols sort alphabetically in id order (DecVocabulary keeps its words sorted, and
    the decoder's output label - 1 is an index into it), three HTK feature files, and a list of extended file names:
    whole files, segments from the start and from the middle, a segment that ends mid-word (no result) and one whose end
    lies beyond the file."""
    from juicer_b200 import synth
    m = synth.make_models(120, 4, sigma_mu=0.9, seed=11, n_gmm_pool=150)
    net = synth.bigram_net(100, 120, k_bigram=5, seed=12)
    words = [f"w{i:04d}" for i in range(100)]
    net.out_names = ["<eps>"] + words
    kw = dict(main_beam=180.0, end_beam=140.0)
    files = synth.make_fixture("h", str(tmp_path), m, net)
    lex = str(tmp_path / "h.lex")
    with open(lex, "w") as f:
        f.write("".join(f"{w} a b\n" for w in words))
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(3)
    paths = []
    for u in range(3):
        x, _ = ps.sample(120, rng)
        p = str(tmp_path / f"utt{u}.htk")
        H.write_htk(p, x)
        paths.append(p)
    p0, p1, p2 = paths
    specs = [p0, f"seg1={p1}[0,90]", f"mid={p2}[20,80]", f"cut={p0}[0,37]", f"late={p1}[25,200]", p2]
    lst = str(tmp_path / "files.lst")
    with open(lst, "w") as f:
        f.write("".join(s + "\n" for s in specs))
    return files, kw, words, lex, lst, specs


class _PortDecoder:
    """decode_batch over the oracle port: lets the CPU suite drive harness.decode_files without a GPU."""

    def __init__(self, port):
        self.p = port

    def decode_batch(self, feats):
        return [self.p.decode(x) for x in feats]


@pytest.mark.parametrize("fmt", ["ref", "trans", "mlf", "xmlf", "verbose"])
def test_output_matches_the_reference_harness(fmt, tmp_path, oracle_port_lib, product_lib):
    """The text harness.decode_files writes equals, byte for byte, what the reference's own DecoderBatchTest::run writes
    for the same file list, network, models and pruning: extended names, frame feeding, result extraction and the
    output format are all the reference's compiled code on that side."""
    from helpers import flat_tables_from_files
    from juicer_b200 import _abi
    from oracle import binding
    if not os.path.exists(binding.HARNESS_SO):
        pytest.skip("oracle/_ref/liboracle_harness.so not built (needs /root/reference)")
    files, kw, words, lex, lst, specs = _harness_fixture(tmp_path)
    want = binding.ref_harness_run(files, lex, lst, str(tmp_path / ("ref_" + fmt)), fmt, **kw)
    tabs, _n, _m = flat_tables_from_files(files)
    port = binding.OraclePort(tabs, _abi.make_cfg(**kw))
    got = H.decode_files(_PortDecoder(port), specs, words, fmt=fmt, expected_dim=39)
    assert got == want
    assert "w0" in want                                      # something was decoded
    port.close()


@pytest.mark.gpu
def test_gpu_output_matches_the_reference_harness(tmp_path, oracle_port_lib, product_lib):
    """The same with the CUDA decoder behind harness.decode_files (xmlf: words, times and per-word scores)."""
    from juicer_b200 import api
    from oracle import binding
    if not os.path.exists(binding.HARNESS_SO):
        pytest.skip("oracle/_ref/liboracle_harness.so did not travel to this box")
    files, kw, words, lex, lst, specs = _harness_fixture(tmp_path)
    net = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
    models = api.HTKFlatModels(files["jmbi"])
    dec = api.WFSTDecoderLite(net, models, 0.0, kw["main_beam"], kw["end_beam"], 0.0, 0, n_lanes=4)
    for fmt in ("xmlf", "verbose"):
        want = binding.ref_harness_run(files, lex, lst, str(tmp_path / ("ref_" + fmt)), fmt, **kw)
        assert H.decode_files(dec, specs, words, fmt=fmt, expected_dim=39) == want
    dec.close()
