#!/bin/bash
# GPU experiment driver: cfg2 bench of build / environment variants, output under gpurun_out/<dir>/
# usage (on the GPU box): bash tools/exp_variants.sh <dir> "name ENV=.. ENV=.." "name2 ..." ...
dir=gpurun_out/$1; shift
mkdir -p $dir
for spec in "$@"; do
  set -- $spec
  name=$1; shift
  env "$@" ABIP_GPU_TUNE_VERBOSE=1 timeout 300 python bench.py --steps 2 --warmup 1 --only none --no-cpu-baseline \
      > $dir/$name.json 2> $dir/$name.err
  python - "$dir/$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    r = d["roofline"]
    print("%-12s it/s %.1f  t %.4f s  e2e %.1f (%.3f s)  frac %.3f  admm %s cg %s  k_bb %.3f ms k_admm %.3f ms" % (
        sys.argv[2], d["value"], d["time_to_1e-4_s"], d["e2e"]["value"], d["e2e"]["time_to_1e-4_s"], r["frac"],
        d["config"].get("admm_iter_per_solve"), d["counters"]["cg_iters"], r["avg_launch_ms"], r["k_admm_iter"]["avg_launch_ms"]))
except Exception as ex:
    print(sys.argv[2], "FAILED", ex)
PY
done
