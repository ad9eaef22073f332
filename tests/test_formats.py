"""On-disk formats next to the path: the JWNT binary network reader
(jgpu_load_jwnt <-> WFSTNetwork::readBinary, src/WFSTNetwork.cpp:1228-1370) against files written by the
reference's own WFSTNetwork::writeBinary (tests/golden/*/*.jwnt, tools/make_golden.py) and against the
reference's own reader (oracle/_ref)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN_CASES, Golden

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
