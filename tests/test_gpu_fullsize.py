"""GPU suite (-m gpu): the BASELINE.json configurations at FULL size against the reference.

configs[2] (c3: 20k-word trigram-shaped network, ~440k states / ~1.8M arcs) and configs[4] (c5: 64k words,
~1.45M states / ~5.9M arcs) are decoded by the unmodified reference objects (oracle/_ref; the plain-C port when
that library is absent) on a bounded sample — two utterances of ~70 frames — at the narrow and the wide end of
the beam sweep, with and without histogram pruning.  Everything must be bit-exact: labels, word-end frames, the
three scores of every word and of the totals, and per frame nActiveInsts / nActiveEmitHyps / nActiveEndHyps /
nEndHypsProcessed and bestEmitScore.  At beam 400 the c5 utterances keep ~690k instances alive in one frame."""
import os

import numpy as np
import pytest

from helpers import bits, flat_tables_from_files, same_result

from juicer_b200 import _abi, api, synth

pytestmark = pytest.mark.gpu

SETTINGS = {
    "c3": [dict(main_beam=250.0), dict(main_beam=400.0), dict(main_beam=250.0, max_hyps=6000)],
    "c5": [dict(main_beam=250.0), dict(main_beam=400.0), dict(main_beam=400.0, max_hyps=6000)],
}


class _Ref:
    """The reference decoder on the fixture: oracle/_ref when it travelled to the box, else the C port."""

    def __init__(self, files, tabs):
        from oracle import binding
        self.files, self.tabs = files, tabs
        self.use_ref = os.path.exists(binding.REF_SO)
        self.o = binding.OracleRef(files, main_beam=250.0) if self.use_ref else None

    def set(self, kw):
        from oracle import binding
        if self.use_ref:
            self.o.set_decoder(**kw)
        else:
            if self.o is not None:
                self.o.close()
            self.o = binding.OraclePort(self.tabs, _abi.make_cfg(**kw))

    def decode(self, x):
        return self.o.decode(x, counters=True)

    def close(self):
        if self.o is not None:
            self.o.close()


@pytest.mark.parametrize("name", ["c3", "c5"])
def test_fullsize_config_matches_reference(name, tmp_path_factory, oracle_port_lib, product_lib, monkeypatch):
    if name == "c3":
        monkeypatch.setenv("JUICER_B200_LAZY", "1")       # c3 also exercises the opt-in per-step scorer at full size
    m, net, tee, _ = synth.named_config(name)
    files = synth.make_fixture(name, str(tmp_path_factory.mktemp(name)), m, net)
    tabs, netl, models = flat_tables_from_files(files)
    ref = _Ref(files, tabs)
    ps = synth.PathSampler(net, m)
    rng = np.random.default_rng(4100)
    xs = [ps.sample(60, rng)[0] for _ in range(2)]
    for kw in SETTINGS[name]:
        ref.set(kw)
        dec = api.WFSTDecoderLite(netl, models, 0.0, kw["main_beam"], 0.0, 0.0, kw.get("max_hyps", 0), n_lanes=2,
                                  frame_stats=True)
        want = [ref.decode(x) for x in xs]
        for u, x in enumerate(xs):
            got = dec.decode(x, lane=u)
            tag = f"{name} {kw} utt{u}"
            same_result(want[u], got, tag)
            cnt, best = dec.frame_stats(u)
            assert np.array_equal(want[u].frame_cnt[:, [0, 1, 2, 4]], cnt), tag
            assert np.array_equal(bits(want[u].frame_best), bits(best)), tag
        for u, r in enumerate(dec.decode_batch(xs)):
            same_result(want[u], r, f"{name} {kw} batch utt{u}")
        dec.close()
    ref.close()
