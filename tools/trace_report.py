"""CTA timeline report of a -DJG_TRACE build:  python tools/trace_report.py trace.bin

Records are {kid, block, smid, aux, t0, t1} (globaltimer ns) written by one thread per CTA at CTA exit.
Per kernel launch (CTAs of one kernel id whose time spans overlap) it prints the launch span, the mean /
p50 / p95 / max CTA duration, how busy the SMs were (sum of CTA time / (span * resident CTAs)), and the
blocks that finish last — the tail the kernel duration is made of."""
import sys
import numpy as np

NAMES = ["gmm", "boundary", "internal", "filter", "expand", "commit_huge", "commit", "expand_r1", "expand_r2"]
dt = np.dtype([("kid", "i4"), ("block", "i4"), ("smid", "i4"), ("aux", "i4"), ("t0", "u8"), ("t1", "u8"), ("mark", "u4", (8,))])


def main():
    r = np.fromfile(sys.argv[1], dtype=dt)
    if len(r) == 0:
        print("empty trace")
        return
    r = r[np.argsort(r["t0"], kind="stable")]
    tmin = int(r["t0"].min())
    # split into launches: same (kid, aux) and a gap-free overlap chain
    launches = []
    cur = [0]
    cur_end = int(r["t1"][0])
    for i in range(1, len(r)):
        same = r["kid"][i] == r["kid"][cur[0]] and r["aux"][i] == r["aux"][cur[0]]
        if same and int(r["t0"][i]) <= cur_end:
            cur.append(i)
            cur_end = max(cur_end, int(r["t1"][i]))
        else:
            launches.append(np.asarray(cur))
            cur = [i]
            cur_end = int(r["t1"][i])
    launches.append(np.asarray(cur))
    agg = {}
    prev_end = None
    for idx in launches:
        x = r[idx]
        t0, t1 = int(x["t0"].min()), int(x["t1"].max())
        dur = (x["t1"] - x["t0"]).astype(np.float64) / 1e3
        span = (t1 - t0) / 1e3
        key = (NAMES[x["kid"][0]] if x["kid"][0] < len(NAMES) else str(x["kid"][0]), int(x["aux"][0]))
        gap = (t0 - prev_end) / 1e3 if prev_end is not None else 0.0
        prev_end = t1
        # time at which 50% / 90% of CTAs were done, relative to launch start
        ends = np.sort((x["t1"] - t0).astype(np.float64) / 1e3)
        a = agg.setdefault(key, [])
        a.append((span, dur.mean(), np.percentile(dur, 50), np.percentile(dur, 95), dur.max(), len(x),
                  ends[len(ends) // 2], ends[int(len(ends) * 0.9)], gap, (x["t0"].max() - t0) / 1e3))
    print(f"{len(r)} CTA records, {len(launches)} launches, {(int(r['t1'].max()) - tmin) / 1e3:.1f} us traced")
    print(f"{'kernel':>16} {'n':>4} {'CTAs':>6} {'span':>8} {'mean':>8} {'p50':>8} {'p95':>8} {'max':>8} {'50%done':>8} {'90%done':>8} {'gap':>6} {'lastStart':>9}")
    tot = 0.0
    for key, v in sorted(agg.items(), key=lambda kv: -sum(t[0] for t in kv[1])):
        m = np.mean(np.asarray(v), axis=0)
        tot += m[0] * len(v)
        print(f"{key[0] + ':' + str(key[1]):>16} {len(v):4d} {m[5]:6.0f} {m[0]:8.1f} {m[1]:8.1f} {m[2]:8.1f} {m[3]:8.1f} {m[4]:8.1f} {m[6]:8.1f} {m[7]:8.1f} {m[8]:6.1f} {m[9]:9.1f}")
    print(f"sum of spans {tot:.1f} us")
    # thread-0 phase marks of the first chunk (median over the CTAs that reached them), us after CTA start
    for kid in np.unique(r["kid"]):
        for aux in np.unique(r["aux"][r["kid"] == kid]):
            x = r[(r["kid"] == kid) & (r["aux"] == aux)]
            mk = x["mark"].astype(np.float64) / 1965.0    # SM cycles -> us at 1965 MHz
            med = [np.median(mk[:, i][mk[:, i] > 0]) if (mk[:, i] > 0).any() else 0.0 for i in range(mk.shape[1])]
            if any(med):
                print(f"   marks {NAMES[kid] if kid < len(NAMES) else kid}:{aux}: " + " ".join(f"{m:6.2f}" for m in med))
    # detail of the longest launch of the two slowest kernels
    for key, v in sorted(agg.items(), key=lambda kv: -sum(t[0] for t in kv[1]))[:4]:
        best = None
        for idx in launches:
            x = r[idx]
            k2 = (NAMES[x["kid"][0]] if x["kid"][0] < len(NAMES) else str(x["kid"][0]), int(x["aux"][0]))
            if k2 == key and (best is None or len(idx) >= len(best)):
                best = idx
        x = r[best]
        t0 = int(x["t0"].min())
        order = np.argsort(x["t1"])[::-1][:6]
        print(f"-- {key}: last CTAs to finish (block, smid, start us, end us)")
        for o in order:
            print(f"     block {x['block'][o]:5d} sm {x['smid'][o]:3d}  {(int(x['t0'][o]) - t0) / 1e3:7.1f} -> {(int(x['t1'][o]) - t0) / 1e3:7.1f}")
        # duration vs block index deciles
        b = x["block"]
        d = (x["t1"] - x["t0"]).astype(np.float64) / 1e3
        o = np.argsort(b)
        parts = np.array_split(d[o], 8)
        print("     mean CTA us by block-index octile:", " ".join(f"{p.mean():.1f}" for p in parts))


if __name__ == "__main__":
    main()
