"""Per-source-line instruction counts and stall samples of an ncu report:
   python tools/ncu_lines.py report.ncu-rep [N] [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
kidx = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split by "Function Name" blocks
blocks = []
cur = None
fname = ""
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        if not blocks or blocks[-1]["name"] != r[1] or blocks[-1].get("closed"):
            pass
        cur = {"name": r[1], "file": fname, "hdr": None, "rows": []}; blocks.append(cur); continue
    if cur is None: continue
    if r[0] == "Line No": cur["hdr"] = r; continue
    if cur["hdr"] and r[0] not in ("", "...") and r[0].isdigit():
        cur["rows"].append((cur["file"], r))
# group blocks of the same kernel launch: consecutive blocks with the same function name
kernels = []
for b in blocks:
    if kernels and kernels[-1][0]["name"] == b["name"] and b["file"] not in [x["file"] for x in kernels[-1]]:
        kernels[-1].append(b)
    else:
        kernels.append([b])
k = kernels[kidx]
print(k[0]["name"])
lines = []
for b in k:
    h = b["hdr"]
    ie = h.index("Instructions Executed"); ws = h.index("Warp Stall Sampling (All Samples)")
    for f, r in b["rows"]:
        try: lines.append((f, int(r[0]), r[1].strip(), float(r[ie] or 0), float(r[ws] or 0)))
        except ValueError: pass
ti = sum(l[3] for l in lines); ts = sum(l[4] for l in lines)
print(f"total warp-instructions {ti:.0f}, stall samples {ts:.0f}")
print("--- by instructions executed")
for l in sorted(lines, key=lambda l: -l[3])[:n]:
    print(f"{l[0]}:{l[1]:4d} {100*l[3]/ti:5.1f}% inst {100*l[4]/max(ts,1):5.1f}% stall  {l[2][:110]}")
print("--- by stall samples")
for l in sorted(lines, key=lambda l: -l[4])[:n]:
    print(f"{l[0]}:{l[1]:4d} {100*l[3]/ti:5.1f}% inst {100*l[4]/max(ts,1):5.1f}% stall  {l[2][:110]}")
