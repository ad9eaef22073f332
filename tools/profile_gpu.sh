#!/bin/bash
# Profiling recipe (B200_PROFILING.md) — run under gpurun; outputs land in gpurun_out/.
# 1) launch list of a (shortened) bench.py run: per-launch device time of every kernel
# 2) ncu --set full on the top kernels
set -x
OUT=gpurun_out
TAG=${1:-r01}
BENCH="python bench.py --workload c3 --utts 64 --lanes 64 --min-frames 100 --max-frames 120 --steps 1 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 4000 -c 1500 --csv --log-file $OUT/launches_${TAG}.csv $BENCH > $OUT/ncu_bench_${TAG}.log 2>&1
for K in k_commit k_internal k_expand k_gmm_scores k_seed; do
  ncu --set full --clock-control none --import-source on -k regex:^${K} -s 300 -c 2 -f -o $OUT/prof_${K}_${TAG} $BENCH > $OUT/ncu_${K}_${TAG}.log 2>&1
done
ls -la $OUT
