"""GPU suite (-m gpu): the decoder has no capacity of its own.  The reference allocates instances, paths and
result records as it goes (src/WFSTDecoderLite.cpp:608-620, 751-805, 279-301); the CUDA path works in
pre-sized arenas, so every arena must either be large enough for anything the network allows or be
re-viewed / grown behind the caller's back.  Also here: exact score ties are broken the same way in every run."""
import numpy as np
import pytest

from helpers import Golden, flat_tables_from_files, same_result

from juicer_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu


def make_decoder(net, models, kw, **extra):
    return api.WFSTDecoderLite(net, models, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                               kw.get("word_beam", 0.0), kw.get("max_hyps", 0), **extra)


def test_overflowing_lane_is_decoded_again_with_a_larger_arena(oracle_port_lib, product_lib):
    """First-pass arenas HALF of what the utterances need on 8 lanes: every utterance overflows its lane, the
    batch call decodes them again with the same pools viewed as 2 lanes x 4 times the arena, and the caller sees
    exactly the reference's results."""
    g = Golden("c2mini")
    tabs, net, models = flat_tables_from_files(g.files)
    ref = make_decoder(net, models, g.kw, n_lanes=1, frame_stats=True)
    ref.decode(g.feats(0))
    cnt, _ = ref.frame_stats(0)
    peak = int(cnt[:, 0].max())
    ref.close()
    assert peak > 256
    dec = make_decoder(net, models, g.kw, n_lanes=8, max_active=peak // 2)
    feats = [g.feats(u % 2) for u in range(12)]
    for rep in range(2):                                   # the second call starts from the enlarged view
        got = dec.decode_batch(feats)
        for u, r in enumerate(got):
            g.check(u % 2, r, what=f"second pass, call {rep}")
    st = dec.stats(-1)
    assert st["n_frames"] == sum(x.shape[0] for x in feats)       # every utterance counted once
    one = dec.decode(g.feats(1)[:20])                      # the streaming interface still sees the (too small) base view
    assert one.status <= -10 and ((-(one.status + 13)) >> 8) & 1
    dec.close()


def test_long_utterance_has_no_word_limit(oracle_port_lib, product_lib):
    """More than 256 words on one best path (the result buffer was hard-wired to 256 records)."""
    from oracle.binding import OraclePort
    m, net, tee, kw = synth.named_config("c1")
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        files = synth.make_fixture("c1", d, m, net)
        tabs, netl, models = flat_tables_from_files(files)
        x, words = synth.PathSampler(net, m).sample(3200, np.random.default_rng(5))
        assert len(words) > 300
        p = OraclePort(tabs, _abi.make_cfg(**kw))
        want = p.decode(x, max_words=4096)
        assert want.status > 300
        dec = make_decoder(netl, models, kw, n_lanes=2, max_words=4096)
        same_result(want, dec.decode(x), "streaming, > 256 words")
        for r in dec.decode_batch([x, x[:40], x]):
            if r.n_frames == x.shape[0]:
                same_result(want, r, "batch, > 256 words")
        dec.close()
        # a caller buffer smaller than the path gets the NEWEST words: the last record carries the totals
        small = make_decoder(netl, models, kw, n_lanes=1, max_words=16)
        r = small.decode(x)
        assert r.status == want.status and len(r.words) == 16
        assert r.labels == want.labels[-16:] and r.times == want.times[-16:]
        assert np.float32(r.words[-1]["ac"]).tobytes() == np.float32(want.words[-1]["ac"]).tobytes()
        small.close(); p.close()


def test_result_word_pool_grows(monkeypatch, oracle_port_lib, product_lib):
    """The words of a batch share one device pool; when it is too small the utterances that did not fit are
    decoded again with a larger one."""
    g = Golden("c1")
    tabs, net, models = flat_tables_from_files(g.files)
    monkeypatch.setenv("JUICER_B200_WORD_POOL", "7")        # less than two results
    dec = make_decoder(net, models, g.kw, n_lanes=4)
    got = dec.decode_batch([g.feats(u % 2) for u in range(10)])
    for u, r in enumerate(got):
        g.check(u % 2, r, what="grown word pool")
    dec.close()


def test_exact_ties_are_broken_the_same_way_every_run(oracle_port_lib, product_lib):
    """`ties` fixture: homophones and equal-weight routes make different arrivals at one state carry bit-identical
    scores in most frames.  The winner is a function of the network (largest arrival id = last arc in file order),
    not of the order in which threads allocated records: 24 lanes decode the same utterances again and again and
    must agree with each other; on this fixture the choice also coincides with the reference's list order
    (src/WFSTDecoderLite.cpp:563-571, instances are prepended to the active list :799-805), so the labels equal the
    golden vectors — in general only the homophone CLASS and the scores are guaranteed to."""
    g = Golden("ties")
    tabs, net, models = flat_tables_from_files(g.files)
    dec = make_decoder(net, models, g.kw, n_lanes=24)
    feats = [g.feats(u % 3) for u in range(72)]
    first = None
    for rep in range(3):
        got = dec.decode_batch(feats)
        for u, r in enumerate(got):
            z = g.z
            want = z[f"labels{u % 3}"].tolist()
            assert r.status == int(z[f"status{u % 3}"])
            assert [synth.TIE_CLASSES[x] for x in r.labels] == [synth.TIE_CLASSES[x] for x in want]
            assert r.times == z[f"times{u % 3}"].tolist()
            assert np.array_equal(np.asarray(r.totals, dtype=np.float32).view(np.uint32), z[f"totals{u % 3}"])
        lab = [r.labels for r in got]
        for u in range(3, 72):
            assert lab[u] == lab[u % 3], f"lanes disagree on a tie (utterance {u}, repetition {rep})"
        first = first or lab
        assert lab == first, f"runs disagree on a tie (repetition {rep})"
    for u in range(3):
        assert first[u] == g.z[f"labels{u}"].tolist()
    dec.close()
