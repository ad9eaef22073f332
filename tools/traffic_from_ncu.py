"""profiles/roofline_traffic.json from an `ncu --set full` report of the dominant kernel on the default bench workload:

    python tools/traffic_from_ncu.py gpurun_out/prof_k_internal_r02.ncu-rep [workload] [kernel key]

traffic = dram__bytes_read.sum + dram__bytes_write.sum of the (first) profiled launch.  The entry records the hash of
the kernel sources it was taken on (bench.csrc_sha): bench.py quotes it only while the sources are unchanged."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def metric_table(rep: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    return hdr, units, rows[2:]


def to_bytes(v: str, unit: str) -> float:
    x = float(v.replace(",", ""))
    u = unit.strip().lower()
    return x * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1.0)


def main():
    rep = sys.argv[1]
    workload = sys.argv[2] if len(sys.argv) > 2 else "c3"
    key = sys.argv[3] if len(sys.argv) > 3 else "k_internal"
    hdr, units, rows = metric_table(rep)
    r = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    rd = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]])
    wr = to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
    dur = r[col["gpu__time_duration.sum"]] + " " + units[col["gpu__time_duration.sum"]]
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        tj = json.load(open(path))
    except Exception:
        tj = {}
    tj[workload] = {key: rd + wr, "csrc_sha": bench.csrc_sha(),
                    "_note": f"ncu --set full --clock-control none: dram__bytes_read.sum + dram__bytes_write.sum ({rd / 1e6:.1f} + {wr / 1e6:.1f} MB) "
                             f"of one launch of {r[col['Kernel Name']]} (launch id {r[col['ID']]}, {dur} under the profiler) in "
                             f"`python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side`; report {os.path.basename(rep)}"}
    json.dump(tj, open(path, "w"), indent=1)
    print(json.dumps(tj[workload], indent=1))


if __name__ == "__main__":
    main()
