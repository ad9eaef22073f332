"""CPU suite: the C-ABI library loads, exports every symbol include/*.h declares, its host
loaders fail loudly on bad input, and there is no CPU compute path."""
import ctypes as C
import os
import re

import pytest

from helpers import Golden, ROOT

from juicer_b200 import _abi, api


def declared_symbols():
    syms = set()
    inc = os.path.join(ROOT, "include")
    for fn in os.listdir(inc):
        if fn.endswith(".h"):
            txt = open(os.path.join(inc, fn)).read()
            txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
            syms |= set(re.findall(r"\b(jgpu_[a-z_0-9]+)\s*\(", txt))
    return sorted(syms)


def test_library_exports_every_declared_symbol(product_lib):
    lib = C.CDLL(product_lib)
    syms = declared_symbols()
    assert len(syms) >= 17
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ but not exported"


def test_struct_sizes_match_header(product_lib):
    # POD layout guards for the ctypes mirror
    assert C.sizeof(_abi.JgpuWord) == 20
    assert C.sizeof(_abi.JgpuCfg) == 48
    assert C.sizeof(_abi.JgpuStats) == 80
    assert C.sizeof(_abi.JgpuResult) == 32


def test_loader_errors_are_loud(tmp_path, product_lib):
    g = Golden("c1")
    with pytest.raises(api.JuicerError, match="cannot open"):
        api.WFSTNetwork(str(tmp_path / "missing.fsm"), g.files["insyms"], g.files["outsyms"])
    with pytest.raises(api.JuicerError, match="cannot open"):
        api.HTKFlatModels(str(tmp_path / "missing.jmbi"))
    bad = tmp_path / "bad.jmbi"
    bad.write_bytes(b"NOPE" + b"\0" * 64)
    with pytest.raises(api.JuicerError, match="not a JMBI"):
        api.HTKFlatModels(str(bad))
    # an arc label without a symbol-table entry is fatal in the reference (isAuxiliary -> error)
    syms = tmp_path / "short.insyms"
    syms.write_text("<eps> 0\nh0 1\n")
    with pytest.raises(api.JuicerError):
        api.WFSTNetwork(g.files["fsm"], str(syms), g.files["outsyms"])


def test_aux_symbols_become_word_end_marker(tmp_path, product_lib):
    """REMOVEBOTH rewrites '#...' labels to wordEndMarker = max label + 1 (WFSTNetwork.cpp:562-565,1438-1455)."""
    (tmp_path / "a.fsm").write_text("0 1 1 1 0.5\n1 0 2 0 0.25\n1 2 0 2\n0 1.5\n")
    (tmp_path / "a.insyms").write_text("<eps> 0\nh0 1\n#aux 2\n")
    (tmp_path / "a.outsyms").write_text("<eps> 0\nW0 1\n#0 2\n")
    n = api.WFSTNetwork(str(tmp_path / "a.fsm"), str(tmp_path / "a.insyms"), str(tmp_path / "a.outsyms"), 2.0, 0.125)
    a = n.arrays()
    assert a["arc_in"].tolist() == [1, 3, 0] and a["arc_out"].tolist() == [1, 0, 3]
    # weight = -w*scale (+ insPenalty when out > 0, tested on the ORIGINAL label)
    assert a["arc_w"].tolist() == [-0.5 * 2.0 + 0.125, -0.25 * 2.0, 0.125]
    assert a["st_final"][0] == -3.0 and a["st_final"][1] == _abi.LOG_ZERO
    assert n.init_state == 0 and a["st_n"].tolist() == [1, 2, 0]


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device the decoder must refuse to exist."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = Golden("c1")
    net = api.WFSTNetwork(g.files["fsm"], g.files["insyms"], g.files["outsyms"])
    models = api.HTKFlatModels(g.files["jmbi"])
    with pytest.raises(api.JuicerError, match="no CUDA device"):
        api.WFSTDecoderLite(net, models, main_beam=200.0)


def test_product_does_not_touch_the_oracle():
    """Nothing under juicer_b200/ or include/ may reference oracle/ (parity claims depend on it)."""
    for base in ("juicer_b200", "include"):
        for dp, _dn, fns in os.walk(os.path.join(ROOT, base)):
            for fn in fns:
                if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                    txt = open(os.path.join(dp, fn), errors="replace").read()
                    assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt \
                        and "juicer_oracle" not in txt, os.path.join(dp, fn)
