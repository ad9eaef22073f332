#!/bin/bash
# quick perf iteration: bench on c3 without the CPU baseline, print the kernel table
TAG=${1:-x}
python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline ${@:2} > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err
python - <<PY
import json
j = json.load(open("gpurun_out/bench_${TAG}.json"))
r = j["roofline"]
print("value %.0f f/s  e2e %.0f f/s  ms/step %.1f  ok_utts %d" % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["utterances_with_result"]))
for k in r["kernel_ms"]:
    n = max(r["kernel_launches"][k], 1)
    print("  %-14s %8.2f ms  %6d launches  %7.2f us/launch  %6.1f GB/s alg" % (k, r["kernel_ms"][k], n, 1e3 * r["kernel_ms"][k] / n, r["kernel_gbs"][k]))
PY
tail -3 gpurun_out/bench_${TAG}.err
