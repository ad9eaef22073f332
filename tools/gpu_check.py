"""Ad-hoc GPU parity probe (development aid; the real parity tests live in tests/)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from juicer_b200 import synth, _abi, api
from oracle.binding import OracleRef, OraclePort

FX = "/tmp/fx"
rng = np.random.default_rng(5)

def cmp(a, b, name):
    ok = True
    if a.status != b.status: print(name, "STATUS", a.status, b.status); ok = False
    if a.labels != b.labels or a.times != b.times: print(name, "WORDS differ"); print(" ", a); print(" ", b); ok = False
    if not np.array_equal(a.totals.view(np.uint32), b.totals.view(np.uint32)):
        print(name, "TOTALS differ", a.totals, b.totals, a.totals - b.totals); ok = False
    return ok

def run(name, m, net, frames, tee=(), n_utts=3, **kw):
    files = synth.make_fixture(name, FX, m, net)
    ps = synth.PathSampler(net, m, tee_hmms=tee)
    xs = [ps.sample(frames, rng)[0] for _ in range(n_utts)]
    o = OracleRef(files, **kw)
    n = api.WFSTNetwork(files["fsm"], files["insyms"], files["outsyms"])
    hm = api.HTKFlatModels(files["jmbi"])
    tabs = _abi.FlatTables(n.arrays(), n.init_state, hm.arrays())
    p = OraclePort(tabs, _abi.make_cfg(**kw))
    dec = api.WFSTDecoderLite(n, hm, kw.get("start_beam", 0.0), kw["main_beam"], kw.get("end_beam", 0.0),
                              kw.get("word_beam", 0.0), kw.get("max_hyps", 0), n_lanes=2, frame_stats=True)
    g_ref = o.gmm_scores(xs[0][:40]); g_gpu = dec.gmm_scores(xs[0][:40])
    nbad = int((g_ref.view(np.uint32) != g_gpu.view(np.uint32)).sum())
    print(f"[{name}] gmm scores: {g_ref.size} values, {nbad} differ, max abs {np.abs(g_ref - g_gpu).max():.3g}")
    allok = True
    for i, x in enumerate(xs):
        r = o.decode(x, counters=True)
        q = p.decode(x, counters=True)
        t0 = time.time(); g = dec.decode(x, lane=i % 2); dt = time.time() - t0
        ok = cmp(r, g, f"{name}/utt{i}")
        cnt, best = dec.frame_stats(i % 2)
        T = x.shape[0]
        if cnt.shape[0] != T: print("frame stats length", cnt.shape, T); ok = False
        else:
            rc = r.frame_cnt[:, [0, 1, 2, 4]]
            if not np.array_equal(rc, cnt):
                bad = np.nonzero((rc != cnt).any(axis=1))[0]
                print(f"  counters differ at {len(bad)} frames, first {bad[:5]}: ref {rc[bad[0]]} gpu {cnt[bad[0]]}"); ok = False
            if not np.array_equal(r.frame_best.view(np.uint32), best.view(np.uint32)):
                bad = np.nonzero(r.frame_best != best)[0]
                print(f"  bestEmit differs at {len(bad)} frames, first {bad[:5]}: {r.frame_best[bad[0]]} vs {best[bad[0]]}"); ok = False
        print(f"[{name}] utt{i} T={T} ok={ok} ref={r.status} words; gpu {T/dt:.0f} f/s (streaming, cold) ref {T/r.seconds:.0f} f/s; stats {dec.stats(i % 2)}")
        allok &= ok
    # batch API
    t0 = time.time(); gb = dec.decode_batch(xs); dt = time.time() - t0
    for i, x in enumerate(xs):
        allok &= cmp(o.decode(x), gb[i], f"{name}/batch{i}")
    print(f"[{name}] batch of {len(xs)}: {sum(x.shape[0] for x in xs)/dt:.0f} f/s, launches={dec.launch_count}, ALL OK={allok}")
    dec.close()
    return allok

ok = True
ok &= run("c1", synth.make_models(10, 1, sigma_mu=2.0, seed=1), synth.digit_loop_net(10), 95, main_beam=200.0)
ok &= run("c1h", synth.make_models(10, 1, sigma_mu=2.0, seed=1), synth.digit_loop_net(10), 95, main_beam=200.0, max_hyps=12)
ok &= run("tee", synth.make_models(4, 2, sigma_mu=2.0, seed=2, with_tee=True), synth.tee_eps_net(4, sp_label=5), 80, tee=[4], main_beam=200.0)
ok &= run("mixed", synth.make_models(40, 3, sigma_mu=1.0, seed=7, with_tee=True, mixed_topology=True, ragged_mix=True),
    synth.bigram_net(30, 40, k_bigram=4, seed=8, sp_label=41), 150, tee=[40], main_beam=150.0, end_beam=100.0, word_beam=80.0, start_beam=120.0, max_hyps=300)
m3 = synth.make_models(2000, 16, sigma_mu=0.8, seed=3); net3 = synth.bigram_net(1000, 2000, k_bigram=8, seed=4)
ok &= run("c2", m3, net3, 300, main_beam=200.0)
ok &= run("c2h", m3, net3, 300, main_beam=200.0, max_hyps=6000, end_beam=150.0)
print("OVERALL", "PASS" if ok else "FAIL")
