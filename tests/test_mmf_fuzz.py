"""Differential tests of the two restatements of the reference's MMF grammar (src/htkparse.l.lpp, src/htkparse.y.ypp):
the product's hand-written scanner + recursive descent (juicer_b200/csrc/host_mmf.cpp) against the oracle's
regex-table scanner + recursive descent (oracle/shim/htkparse_rd.cpp, linked into oracle/_ref in place of the bison
output that cannot be generated here).  Valid files: identical tables through the reference's own load semantics.
Mutated files: the same accept / reject decision at the parse stage.  Needs oracle/_ref (skipped without it)."""
import ctypes as C
import os
import random
import re

import numpy as np
import pytest

from helpers import same_model_tables

from juicer_b200 import api
from oracle import binding

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference") and not os.path.exists(binding.REF_SO),
                                reason="oracle/_ref is not built (no /root/reference here)")

KW = ["BeginHMM", "EndHMM", "NumStates", "State", "NumMixes", "Mixture", "Mean", "Variance", "GConst", "TransP",
      "VecSize", "StreamInfo", "DiagC", "NullD"]


def _fmt(rng: random.Random, x: float) -> str:
    """One of the spellings the RE / INT token classes accept for the same float32 value."""
    x = float(np.float32(x))
    style = rng.randrange(6)
    if style == 0 and x == int(x) and abs(x) < 1000:
        return str(int(x))                                  # INTEGER token inside a real vector
    if style == 1:
        return "%.9e" % x
    if style == 2:
        return ("%.9E" % x).replace("E", "e" if rng.random() < 0.5 else "E")
    if style == 3 and x > 0:
        return "+%.9g" % x if "e" not in "%.9g" % x or True else "%.9g" % x
    if style == 4 and 0 < abs(x) < 1:
        s = "%.9f" % abs(x)
        return ("-" if x < 0 else "") + s[1:]               # ".5" form (no leading zero)
    return "%.9g" % x if "inf" not in "%.9g" % x else "%.9e" % x


def _kw(rng: random.Random, k: str) -> str:
    return "<" + rng.choice([k, k.upper(), k.lower()]) + ">"


def _ws(rng: random.Random) -> str:
    return rng.choice([" ", "\n", "\t", "  ", " \r\n", "\n\n "])


def random_mmf(seed: int) -> str:
    rng = random.Random(seed)
    nrng = np.random.default_rng(seed)
    D = rng.choice([1, 2, 3, 5])
    out = ["~o"]
    opts = [f"{_kw(rng, 'VecSize')} {D}", f"{_kw(rng, 'StreamInfo')} 1 {D}", _kw(rng, "DiagC"), _kw(rng, "NullD"),
            rng.choice(["<MFCC_E_D_A>", "<USER>", "<PLP_0_D>", "<mfcc_d_a_z>", "<LPCepstra_E>"]),
            rng.choice(['<HmmSetId> set_1', '<HMMSETID> "set_1"'])]
    rng.shuffle(opts)
    out += opts[: rng.randrange(2, len(opts) + 1)]
    if not any("ecsize" in o.lower() for o in out):
        out.append(f"{_kw(rng, 'VecSize')} {D}")

    def vec(n):
        return " ".join(_fmt(rng, v) for v in nrng.standard_normal(n) * 3)

    def var(n):
        return " ".join(_fmt(rng, v) for v in nrng.uniform(0.25, 4.0, n))

    def gmm():
        n = rng.randrange(1, 4)
        w = nrng.dirichlet(np.full(n, 3.0)).astype(np.float32)
        parts = []
        if n > 1 or rng.random() < 0.3:
            parts.append(f"{_kw(rng, 'NumMixes')} {n if rng.random() < 0.8 else n + 2}")   # the count is never checked
        for c in range(n):
            if n > 1:
                parts.append(f"{_kw(rng, 'Mixture')} {c + 1} %.9e" % float(w[c]))
            elif parts and rng.random() < 0.5:
                parts.append(f"{_kw(rng, 'Mixture')} 1 1.0")
            parts.append(f"{_kw(rng, 'Mean')} {D} {vec(D)}")
            parts.append(f"{_kw(rng, 'Variance')} {D} {var(D)}")
            if rng.random() < 0.6:
                parts.append(f"{_kw(rng, 'GConst')} {_fmt(rng, float(nrng.normal(50, 10)))}")
        return parts

    def transp(n, tee):
        m = np.zeros((n, n), dtype=np.float32)
        m[0, 1] = 1.0
        if tee:
            m[0, 1], m[0, n - 1] = 0.75, 0.25
        for i in range(1, n - 1):
            m[i, i], m[i, i + 1] = 0.6, 0.4
            if i + 2 <= n - 1 and rng.random() < 0.4:
                m[i, i], m[i, i + 1], m[i, i + 2] = 0.5, 0.3, 0.2
        return [f"{_kw(rng, 'TransP')} {n}"] + [" ".join(_fmt(rng, v) for v in row) for row in m]

    n_sh_t = rng.randrange(0, 3)
    sh_t = []
    for k in range(n_sh_t):
        n = rng.choice([3, 4, 5])
        sh_t.append((f"T_{k}", n))
        out += [f'~t "T_{k}"'] + transp(n, rng.random() < 0.3)
    sh_s = []
    for k in range(rng.randrange(0, 4)):
        sh_s.append(f"s{k}:a-b+c[2]")
        out += [f'~s "{sh_s[-1]}"'] + gmm()
    if rng.random() < 0.4:
        out += ['~v "varFloor1"', f"{_kw(rng, 'Variance')} {D} {var(D)}"]
    for h in range(rng.randrange(1, 6)):
        use_t = rng.choice(sh_t) if sh_t and rng.random() < 0.6 else None
        n = use_t[1] if use_t else rng.choice([3, 4, 5, 6])
        out += [f'~h "m{h}-x+y"', _kw(rng, "BeginHMM"), f"{_kw(rng, 'NumStates')} {n}"]
        if rng.random() < 0.2:
            out += ["~o", f"{_kw(rng, 'VecSize')} {D}"]
        for s in range(2, n):
            out.append(f"{_kw(rng, 'State')} {s if rng.random() < 0.9 else 7}")        # the number is ignored
            if sh_s and rng.random() < 0.4:
                out.append(f'~s "{rng.choice(sh_s)}"')
            else:
                out += gmm()
        out += [f'~t "{use_t[0]}"'] if use_t else transp(n, n == 3 and rng.random() < 0.5)
        out.append(_kw(rng, "EndHMM"))
    return "".join(tok + _ws(rng) for tok in out)


@pytest.fixture(scope="module")
def ref_lib(product_lib):
    binding.build(ref=True, port=False)
    if not os.path.exists(binding.REF_SO):
        pytest.skip("oracle/_ref is not built")
    lib = C.CDLL(binding.REF_SO)
    lib.oref_mmf_parse_only.argtypes = [C.c_char_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("seed", range(24))
def test_random_valid_mmf_same_tables(seed, tmp_path, ref_lib):
    """Random model sets in random spellings (keyword case, whitespace, integer / '.5' / exponent number forms,
    macros in any legal order): the reference's loader on the oracle's parser and jgpu_load_mmf agree bit for bit."""
    p = tmp_path / "r.mmf"
    p.write_text(random_mmf(seed))
    assert ref_lib.oref_mmf_parse_only(str(p).encode(), None) == 0
    for remove_tee in (False, True):
        ref = binding.RefModels(str(p), remove_tee=remove_tee)
        mine = api.HTKFlatModels.from_mmf(str(p), remove_tee)
        same_model_tables(ref.dump_models(), mine.arrays(), f"seed {seed} remove_tee={remove_tee}")
        ref.close()


def _mutate(rng: random.Random, text: str) -> str:
    toks = re.findall(r"\S+|\s+", text)
    idx = [i for i, t in enumerate(toks) if not t.isspace()]
    i = rng.choice(idx)
    kind = rng.randrange(8)
    if kind == 0:
        del toks[i]
    elif kind == 1:
        toks.insert(i, toks[i] + " ")
    elif kind == 2:
        j = rng.choice(idx)
        toks[i], toks[j] = toks[j], toks[i]
    elif kind == 3:
        toks[i] = rng.choice(["1", "-3", "0.5", "1e", "e5", "+", "-", ".", "<Foo>", "~u", '"q"', "abc", "1.5.2", "<Mean>", "<GConst>",
                              "~s", '~s "x"', "<MEAN> 2", "7e-2", "<TMix> p", "~o", "<NumMixes> 2", "<Mixture> 2 0.5"])
    elif kind == 4:
        toks[i] = toks[i].swapcase()
    elif kind == 5:
        toks[i] = toks[i][: max(1, len(toks[i]) // 2)]
    elif kind == 6:
        toks[i] = toks[i].replace(">", "> ", 1) if ">" in toks[i] else toks[i] + "x"
    else:
        toks[i] = rng.choice(["<", ">", "~", '"', "%", "(", "!"]) + toks[i]
    return "".join(toks)


def test_mutated_mmf_same_parse_decision(tmp_path, ref_lib):
    """400 single-token mutations of random files: both parsers accept (and then the tables are identical), or both
    reject at the parse stage.  (What
    the reference rejects later, in initFromHTKParseResult, ends its process; for those files only the product's
    message class is checked.)"""
    rng = random.Random(2024)
    n_accept = n_reject = n_semantic = 0
    for k in range(400):
        text = _mutate(rng, random_mmf(rng.randrange(24)))
        p = tmp_path / f"m{k}.mmf"
        p.write_text(text)
        rc = ref_lib.oref_mmf_parse_only(str(p).encode(), None)
        try:
            m = api.HTKFlatModels.from_mmf(str(p))
            mine = "ok"
            if rc == 0:                                   # accepted by both: the mutation must also MEAN the same
                ref = binding.RefModels(str(p))
                same_model_tables(ref.dump_models(), m.arrays(), f"mutation {k}")
                ref.close()
            del m
        except api.JuicerError as e:
            msg = str(e)
            mine = "parse" if ("HTKPARSE" in msg or "syntax error" in msg or "not supported" in msg) else "semantic"
        if rc == 0:
            assert mine in ("ok", "semantic"), (k, rc, mine, text[:300])
            n_accept += mine == "ok"
            n_semantic += mine == "semantic"
        else:
            assert mine == "parse", (k, rc, mine, text[:300])
            n_reject += 1
    assert n_accept > 20 and n_reject > 100, (n_accept, n_reject, n_semantic)
