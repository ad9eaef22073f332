#!/bin/bash
# A/B of the scorer/search overlap (JUICER_B200_OVERLAP): parity under the persistent scorer first, then bench lines.
#   tools/overlap_gpu.sh "name:ENV=.. ENV=.. ;name2:..."  [bench args]
set -u
JUICER_B200_OVERLAP=1 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "golden or mmf or ragged or c2_single or c3_scaled" 2>&1 | tail -3
IFS=';' read -ra VARS <<< "$1"
for v in "${VARS[@]}"; do
  name=${v%%:*}; envs=${v#*:}
  env $envs python bench.py --workload c3 --steps 2 --warmup 1 --no-cpu-baseline ${@:2} > gpurun_out/bench_ovl_$name.json 2> gpurun_out/bench_ovl_$name.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/bench_ovl_$name.json")); r = j["roofline"]
    print("$name [$envs]: value %.0f f/s e2e %.0f ms/step %.1f ok %d | " % (j["value"], j["e2e"]["value"], j["ms_per_step"], j["e2e"]["utterances_with_result"]) +
          " ".join("%s %.1f" % (k.replace("k_", ""), 1e3 * r["kernel_ms"][k] / max(r["kernel_launches"][k], 1)) for k in r["kernel_ms"]))
except Exception as e:
    print("$name failed:", e); print(open("gpurun_out/bench_ovl_$name.err").read()[-1500:])
PY
done
