"""Summarises gpurun_out/ ncu artefacts into profiles/ (tracked).
   python tools/summarize_profiles.py r01
Writes profiles/<tag>_launches.md (per-kernel share of the step from the ncu launch list)
and profiles/<tag>_<kernel>.md (selected --set full metrics + top stall locations)."""
import csv, glob, os, subprocess, sys, collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, "gpurun_out"); P = os.path.join(ROOT, "profiles"); os.makedirs(P, exist_ok=True)

def launches():
    fn = os.path.join(G, f"launches_{tag}.csv")
    if not os.path.exists(fn): return
    rows = [r for r in csv.reader(open(fn)) if r and r[0].isdigit()]
    hdr = None
    for r in csv.reader(open(fn)):
        if r and r[0] == "ID": hdr = r; break
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    ui = hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1000.0 if r[ui] == "ns" else v * (1000.0 if r[ui] == "ms" else 1.0)   # -> us
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(P, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list ({tag}): gpu__time_duration.sum per launch, --clock-control none\n\n")
        f.write("Command: see tools/profile_gpu.sh (shortened bench.py run on c3, 256 lanes, 256 utterances of 100-120 frames). Per-launch times are cold-cache\n"
                "and serialised: compare SHARES with bench.py's `roofline.kernel_share_of_step`, not absolutes.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {k} | {n} | {t:.1f} | {t/n:.2f} | {100*t/tot:.1f}% |\n")
    print(open(os.path.join(P, f"{tag}_launches.md")).read())

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_pipe_fma.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]

def kernel(rep):
    name = os.path.basename(rep).replace("prof_", "").replace(f"_{tag}.ncu-rep", "")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3: return
    hdr, units = rows[0], rows[1]
    with open(os.path.join(P, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# ncu --set full --clock-control none: {name} ({tag})\n\n")
        for r in rows[2:]:
            f.write(f"## launch id {r[0]}: `{r[hdr.index('Kernel Name')]}` grid {r[hdr.index('Grid Size')] if 'Grid Size' in hdr else ''}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in hdr:
                    f.write(f"| {w} | {r[hdr.index(w)]} | {units[hdr.index(w)]} |\n")
            f.write("\n")
        hot = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_hot.py"), rep, "14"], capture_output=True, text=True).stdout
        f.write("## top warp-stall locations (SASS, first captured launch)\n\n```\n" + hot + "```\n")
    print("wrote", name)

launches()
for rep in sorted(glob.glob(os.path.join(G, f"prof_*_{tag}.ncu-rep"))):
    kernel(rep)
