"""The exactness argument behind jg_mix_logsum's prefilter (juicer_b200/csrc/jgpu_gmm.cuh, off by default): a logAdd step
whose component lies more than the threshold below the running maximum of the EARLIER components cannot change the
accumulator, so skipping it leaves the sequence of roundings of HTKFlatModels::logAdd (src/HTKFlatModels.cpp:266-293)
untouched.  Checked here on the CPU with a float32 / float64 restatement of logAdd; the CUDA side is covered by the
-m gpu parity tests whichever way the switch is set."""
import math
import struct

import numpy as np

LOG_ZERO = np.float32(-3.4028234663852886e38)          # -FLT_MAX, the reference's LOG_ZERO
THR_F32 = struct.unpack("<f", struct.pack("<I", 0xC1935C28))[0]   # smallest float above the double threshold -18.42


def log_add(x: np.float32, y: np.float32) -> np.float32:
    if x < y:
        x, y = y, x
    with np.errstate(over="ignore"):
        diff = np.float32(y - x)                       # float subtraction, rounded once
    if float(diff) < -18.42:                           # compared in double
        return x
    return np.float32(float(x) + math.log(1.0 + math.exp(float(diff))))


def chain(vals) -> np.float32:
    lp = LOG_ZERO
    for v in vals:
        lp = log_add(lp, v)
    return lp


def chain_prefiltered(vals) -> np.float32:
    m = LOG_ZERO
    todo = []
    for k, v in enumerate(vals):
        with np.errstate(over="ignore"):
            if not (np.float32(v - m) < np.float32(THR_F32)):
                todo.append(k)
        m = max(m, v)
    lp = LOG_ZERO
    for k in todo:
        lp = log_add(lp, vals[k])
    return lp


def test_float_threshold_is_the_double_threshold():
    f = np.float32(THR_F32)
    below = np.nextafter(f, np.float32(-100.0))
    assert not (float(f) < -18.42) and float(below) < -18.42


def test_prefiltered_chain_has_the_same_bits():
    rng = np.random.default_rng(5)
    for trial in range(4000):
        n = int(rng.integers(1, 33))
        spread = float(rng.choice([0.5, 5.0, 18.0, 19.0, 40.0, 300.0]))
        base = float(rng.normal(-60.0, 30.0))
        vals = (base + spread * rng.standard_normal(n)).astype(np.float32)
        if trial % 7 == 0:                             # ties, exact threshold distances, a component at LOG_ZERO
            vals[rng.integers(0, n)] = vals[0]
            vals[rng.integers(0, n)] = np.float32(float(vals[0]) - 18.42)
        if trial % 11 == 0:
            vals[rng.integers(0, n)] = LOG_ZERO
        a, b = chain(list(vals)), chain_prefiltered(list(vals))
        assert a.tobytes() == b.tobytes(), (trial, vals, a, b)
