"""Times k_gmm_scores alone on the c3 acoustic models (6000 tied 16-mix GMMs, D=39):
   JUICER_B200_LIB=... python tools/gmm_bench.py [rows]"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from juicer_b200 import api, synth

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = synth.make_models(4000, 16, sigma_mu=1.3, seed=21, n_gmm_pool=6000)
net = synth.digit_loop_net(10)
d = tempfile.mkdtemp()
files = synth.make_fixture("g", d, m, net)
network = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
models = api.HTKFlatModels(files["jmbi"])
dec = api.WFSTDecoderLite(network, models, 0.0, 200.0, 0.0, 0.0, 0, n_lanes=1, device=0)
x = np.random.default_rng(0).normal(size=(rows, m.dim)).astype(np.float32)
dec.gmm_scores(x[:1024])
dec.profile(True)
sc = dec.gmm_scores(x)
p = dec.profile_read()["k_gmm_scores"]
gauss = rows * 6000 * 16
print(f"{os.environ.get('JUICER_B200_LIB', 'default'):45s} {p['ms']:8.3f} ms / {p['launches']} launches  "
      f"{gauss * 39 * 4 / p['ms'] / 1e9:6.2f} T fp32-op/s  checksum {float(np.abs(sc).sum()):.6e}")
dec.close()
