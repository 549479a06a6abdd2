"""Benchmark-scale fixtures: runs the UNMODIFIED reference ABIP-LP solver (oracle/_ref, OpenMP build:
only the SpMV loop is parallel and its results are bit-identical to the serial build) on
  * cfg2 at scale 0.1, 0.25 and 1.0 (BASELINE.json configs[1], the bench workload: m=200k n=1M nnz=5M),
  * cfg4 at scale 0.05 (configs[3] family),
  * the first 64 problems of cfg5 (configs[4]).
Takes ~20-30 minutes of CPU; run in the build container:

    make -C oracle lp && OMP_NUM_THREADS=8 python tests/golden/make_golden_large.py

Writes tests/golden/lp_golden_large.json and tests/golden/cfg5_golden.json; nothing under /root/reference is read at
test time.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from abip_b200 import problems  # noqa: E402
from oracle import ref_lp  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
WHICH = "indirect_omp" if ref_lp.available("indirect_omp") else "indirect"


def record(p, r):
    return {"m": p.m, "n": p.n, "nnz": p.nnz, "status": r["status"], "status_val": r["status_val"],
            "ipm_iter": r["ipm_iter"], "admm_iter": r["admm_iter"], "pobj": r["pobj"], "dobj": r["dobj"],
            "res_pri": r["res_pri"], "res_dual": r["res_dual"], "rel_gap": r["rel_gap"],
            "solve_time_ms": r["solve_time_ms"],
            "x_head": r["x"][:16].tolist(), "y_head": r["y"][:16].tolist(), "s_head": r["s"][:16].tolist(),
            "x_norm": float(np.linalg.norm(r["x"])), "y_norm": float(np.linalg.norm(r["y"])),
            "s_norm": float(np.linalg.norm(r["s"]))}


def main():
    only = sys.argv[1:]  # optional list of case names
    path = os.path.join(HERE, "lp_golden_large.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    cases = {
        "cfg2_scale0.1": lambda: problems.cfg2(scale=0.1),
        "cfg2_scale0.25": lambda: problems.cfg2(scale=0.25),
        "cfg4_scale0.05": lambda: problems.cfg4(scale=0.05),
        "cfg2_full": lambda: problems.cfg2(scale=1.0),
    }
    for name, gen in cases.items():
        if only and name not in only:
            continue
        p = gen()
        t = time.time()
        r = ref_lp.solve(p, which=WHICH, eps=1e-4)
        out[name] = record(p, r)
        out[name]["threads"] = int(os.environ.get("OMP_NUM_THREADS", "1")) if WHICH == "indirect_omp" else 1
        print(name, out[name]["status"], out[name]["ipm_iter"], out[name]["admm_iter"], out[name]["pobj"],
              "%.1f s" % (time.time() - t), flush=True)
        json.dump(out, open(path, "w"), indent=1)
    if not only or "cfg5" in only:
        batch = []
        for i, p in enumerate(problems.cfg5_batch(64)):
            r = ref_lp.solve(p, which="indirect", eps=1e-4)
            batch.append({"seed": 5000 + i, "status": r["status"], "ipm_iter": r["ipm_iter"], "admm_iter": r["admm_iter"],
                          "pobj": r["pobj"], "dobj": r["dobj"], "x_norm": float(np.linalg.norm(r["x"]))})
        json.dump(batch, open(os.path.join(HERE, "cfg5_golden.json"), "w"), indent=1)
        print("cfg5: 64 problems", flush=True)


if __name__ == "__main__":
    main()
