#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun; outputs land in gpurun_out/.
#   tools/profile_gpu.sh TAG [launches] [kernel-regex[:skip[:count]] ...]
# 1) "launches": launch list of a (shortened) bench.py run: per-launch device time of every kernel
# 2) ncu --set full on the kernels named (function-name regexes, e.g. '^k_walk$:300:4' '^k_gmm:2:1'), on the
#    DEFAULT bench workload (c3, 512 utterances, 256 lanes), so that dram__bytes of a launch is the `traffic`
#    of the bench line's roofline (tools/traffic_from_ncu.py turns the report into profiles/roofline_traffic.json)
set -x
OUT=gpurun_out
TAG=${1:-r02}; shift
SHORT="python bench.py --workload c3 --utts 256 --lanes 256 --min-frames 100 --max-frames 120 --steps 1 --warmup 1 --no-cpu-baseline --no-side"
FULL="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-side"
for K in "$@"; do
  if [ "$K" == "launches" ]; then
    ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1500 --csv --log-file $OUT/launches_${TAG}.csv $SHORT > $OUT/ncu_bench_${TAG}.log 2>&1
  else
    IFS=: read -r RE SKIP CNT <<< "$K"
    N=$(echo "$RE" | tr -cd 'a-z_')
    ncu --set full --clock-control none --import-source on -k "regex:$RE" -s ${SKIP:-300} -c ${CNT:-2} -f -o $OUT/prof_${N}_${TAG} $FULL > $OUT/ncu_${N}_${TAG}.log 2>&1
  fi
done
ls -la $OUT
