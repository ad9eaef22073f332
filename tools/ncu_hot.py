"""Top stall locations of an ncu report (SASS view):  python tools/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[0]
h = b["hdr"]; k = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
tot = sum(float(r[k] or 0) for r in b["rows"])
print(b["name"], "total samples", tot)
idx = sorted(range(len(b["rows"])), key=lambda i: -float(b["rows"][i][k] or 0))[:n]
for i in sorted(idx):
    r = b["rows"][i]
    print(f"{i:4d} {float(r[k]):8.0f} {100*float(r[k])/tot:5.1f}%  exec={r[ie]:>8}  {r[1].strip()}")
