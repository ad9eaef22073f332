"""GPU suite: the C++ host adapter GpuWFSTDecoder behind the reference's own Juicer::IDecoder
interface, against WFSTDecoderLite built from the SAME WFSTNetwork* / IModels* objects
(oracle/_ref/dropin_test, compiled against the reference headers where /root/reference exists;
the binary travels to the GPU box)."""
import os
import subprocess

import numpy as np
import pytest

from helpers import Golden

from oracle import binding

pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not os.path.exists(binding.DROPIN_BIN), reason="oracle/_ref/dropin_test not built (needs /root/reference)")
@pytest.mark.parametrize("case,utt", [("c1", 0), ("tee", 1), ("mixed", 2), ("c2mini", 0), ("c2mini", 1), ("nolabel", 0)])
def test_cpp_adapter_is_a_drop_in_for_wfstdecoderlite(case, utt, tmp_path):
    g = Golden(case)
    f = tmp_path / "feats.f32"
    g.feats(utt).astype(np.float32).tofile(str(f))
    kw = g.kw
    cmd = [binding.DROPIN_BIN, g.files["jmbi"], g.files["fsm"], g.files["insyms"], g.files["outsyms"], str(f),
           str(kw["main_beam"]), str(kw.get("end_beam", 0.0)), str(kw.get("word_beam", 0.0)),
           str(kw.get("start_beam", 0.0)), str(kw.get("max_hyps", 0))]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0 and "DROPIN PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    want = int(g.z[f"status{utt}"])
    assert f"reference {want} words, gpu {want} words" in p.stdout


@pytest.mark.skipif(not os.path.exists(binding.DROPIN_BIN), reason="oracle/_ref/dropin_test not built (needs /root/reference)")
@pytest.mark.parametrize("case,utt", [("mixed", 0), ("c2mini", 0), ("tee", 0)])
def test_cpp_adapter_partial_decoding_switch(case, utt, tmp_path):
    """PartialTraceInterval (PARTIAL_DECODING, src/WFSTDecoderLite.cpp:117, 247-257): with the switch on both decoders end
    an utterance with the line "Partial paths recovered at frames: ..."; for a decoded utterance it lists the end frame
    of every word of the best path, and the adapter prints what the reference prints."""
    g = Golden(case)
    f = tmp_path / "feats.f32"
    g.feats(utt).astype(np.float32).tofile(str(f))
    kw = g.kw
    cmd = [binding.DROPIN_BIN, g.files["jmbi"], g.files["fsm"], g.files["insyms"], g.files["outsyms"], str(f),
           str(kw["main_beam"]), str(kw.get("end_beam", 0.0)), str(kw.get("word_beam", 0.0)),
           str(kw.get("start_beam", 0.0)), str(kw.get("max_hyps", 0))]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, DROPIN_PARTIAL="20"))
    assert p.returncode == 0 and "DROPIN PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.startswith("Partial paths recovered at frames:")]
    assert len(lines) == 4                                   # (gpu, reference) x 2 repetitions
    assert lines[0] == lines[1] == lines[2] == lines[3]
    assert [int(x) for x in lines[0].split(":")[1].split()] == g.z[f"times{utt}"].tolist()
