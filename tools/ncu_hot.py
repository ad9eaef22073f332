"""Top stall locations of an ncu report:  python tools/ncu_hot.py report.ncu-rep [N] [sass|cuda]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
view = sys.argv[3] if len(sys.argv) > 3 else "sass"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", view], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}; blocks.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
b = blocks[int(sys.argv[4]) if len(sys.argv) > 4 else 0]
h = b["hdr"]; k = h.index("Warp Stall Sampling (All Samples)"); ie = h.index("Instructions Executed")
src = h.index("Source")
tot = sum(float(r[k] or 0) for r in b["rows"] if len(r) > k)
print(b["name"], "total samples", tot)
rows_ok = [r for r in b["rows"] if len(r) > k]
idx = sorted(range(len(rows_ok)), key=lambda i: -float(rows_ok[i][k] or 0))[:n]
for i in sorted(idx):
    r = rows_ok[i]
    print(f"{i:4d} {float(r[k] or 0):8.0f} {100*float(r[k] or 0)/tot:5.1f}%  exec={r[ie]:>9}  {r[src].strip()[:120]}")
