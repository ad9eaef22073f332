#!/bin/bash
# quick A/B of library variants: tools/variants_gpu.sh "A B C" [bench args]
for v in $1; do
  JUICER_B200_LIB=juicer_b200/libjuicer_b200_$v.so python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline --no-side ${@:2} > gpurun_out/bench_var_$v.json 2> gpurun_out/bench_var_$v.err
  python - <<PY
import json
j = json.load(open("gpurun_out/bench_var_$v.json")); r = j["roofline"]
print("$v: value %.0f f/s ms/step %.1f ok %d | " % (j["value"], j["ms_per_step"], j["e2e"]["utterances_with_result"]) +
      " ".join("%s %.1f" % (k.replace("k_", ""), 1e3 * r["kernel_ms"][k] / max(r["kernel_launches"][k], 1)) for k in r["kernel_ms"]))
PY
done
