#!/bin/bash
# CTA timeline of a few frame steps (needs `python juicer_b200/build.py --trace`):  tools/trace_gpu.sh TAG [bench args]
TAG=${1:-x}; shift
JUICER_B200_LIB=juicer_b200/libjuicer_b200_trace.so JUICER_B200_TRACE=gpurun_out/trace_${TAG}.bin \
  python bench.py --workload c3 --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/trace_bench_${TAG}.json 2> gpurun_out/trace_bench_${TAG}.err
python tools/trace_report.py gpurun_out/trace_${TAG}.bin > gpurun_out/trace_${TAG}.txt 2>&1
tail -40 gpurun_out/trace_${TAG}.txt
