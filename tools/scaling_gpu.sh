#!/bin/bash
# One point of a scaling table, run under `gpurun --gpus N`:
#   tools/scaling_gpu.sh N strong|weak TAG [extra bench.py args]
# strong = BASELINE configs[3]: 10 000 utterances of the c3 workload behind one shared utterance queue (whole-utterance
# work stealing), whatever N;  weak = the driver's own scaling run (N x 512 utterances behind the shared queue).
N=$1; MODE=$2; TAG=$3; shift 3
OUT=gpurun_out
mkdir -p $OUT
ARGS="--gpus $N --scaling $MODE --steps 2 --warmup 1 --no-cpu-baseline --no-side $@"
if [ "$N" == "1" ]; then
  python bench.py $ARGS > $OUT/scale_${TAG}_${MODE}_n$N.json 2> $OUT/scale_${TAG}_${MODE}_n$N.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py $ARGS \
      > $OUT/scale_${TAG}_${MODE}_n$N.json 2> $OUT/scale_${TAG}_${MODE}_n$N.err
fi
tail -c 400 $OUT/scale_${TAG}_${MODE}_n$N.err
python - <<PY
import json
try:
    d = json.load(open("$OUT/scale_${TAG}_${MODE}_n$N.json"))
    print("N=$N $MODE: value %.0f f/s, e2e %.0f f/s, ms/step %.1f, ok %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["utterances_with_result"]), d.get("queue"))
except Exception as e:
    print("no result line:", e)
PY
