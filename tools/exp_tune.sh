#!/bin/bash
# GPU experiment: measured balance (tune_balance) on/off and its knobs at cfg2; output under gpurun_out/tune/
# usage (on the GPU box): bash tools/exp_tune.sh
mkdir -p gpurun_out/tune
run() {  # name, env...
  name=$1; shift
  env "$@" ABIP_GPU_TUNE_VERBOSE=1 ABIP_GPU_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 1 --only none --no-cpu-baseline \
      > gpurun_out/tune/$name.json 2> gpurun_out/tune/$name.err
  python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/tune/{n}.json"))
    print(n, "it/s %.1f  t %.4f s  e2e %.1f  frac %.3f  admm %s" % (d["value"], d["time_to_1e-4_s"], d["e2e"]["value"], d["roofline"]["frac"], d["config"]["admm_iter_per_solve"]))
except Exception as ex:
    print(n, "FAILED", ex)
PY
  grep "balance round" gpurun_out/tune/$name.err | head -12
}
run off ABIP_GPU_TUNE=0
run r3 ABIP_GPU_TUNE_ROUNDS=3
run r3b ABIP_GPU_TUNE_ROUNDS=3
run r6 ABIP_GPU_TUNE_ROUNDS=6
run r6d5 ABIP_GPU_TUNE_ROUNDS=6 ABIP_GPU_TUNE_DAMP=0.5
run r4d10 ABIP_GPU_TUNE_ROUNDS=4 ABIP_GPU_TUNE_DAMP=1.0
run r4reps8 ABIP_GPU_TUNE_ROUNDS=4 ABIP_GPU_TUNE_REPS=8
