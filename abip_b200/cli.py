"""Command-line entry: `python -m abip_b200.cli problem.mps [...]`.

The reference ships no CLI ("C interface is planned", README.md:37); its benchmark scripts call a binary as
`bin/abip-indirect <mps> <timelimit> 100000 10000000 0 1e-10 1e-<precision> 5 1 <out-prefix>`
(scripts/bench-lp/run_all_abip-binary-nobar-indirect.sh:50).  That positional form is accepted (time limit, maximum IPM
iterations, maximum ADMM iterations, tolerance at position 7, output prefix last; the other positions are read and
ignored) next to named options.  The file goes through abip_b200.mps (mpsread + preprocess.m restated), the standard
form through abip_b200.api.abip -- i.e. through the C ABI of libabip_gpu.so; there is no CPU fallback.

Writes <out>.json (status, iterations, objectives incl. the objective constant, residuals, times; with the keys that
scripts/bench-lp/analyze_abip.py reads for both of the reference's result schemas) and <out>.sol (x in the variables
of the file, one value per line).
"""
from __future__ import annotations

import argparse
import json
import sys
import time


def parse_args(argv):
    ap = argparse.ArgumentParser(prog="abip-gpu", description=__doc__.split("\n\n")[0])
    ap.add_argument("mps")
    ap.add_argument("legacy", nargs="*", help="positional form of the reference's benchmark scripts")
    ap.add_argument("--time-limit", type=float, default=None, help="seconds")
    ap.add_argument("--max-ipm-iters", type=int, default=None)
    ap.add_argument("--max-admm-iters", type=int, default=None)
    ap.add_argument("--tol", type=float, default=1e-4)
    ap.add_argument("--out", default=None, help="output prefix (default: the file name without .mps)")
    ap.add_argument("--verbose", type=int, default=0)
    a = ap.parse_args(argv)
    L = a.legacy
    if L:
        if len(L) != 9:
            ap.error("the positional form takes 9 values after the file: timelimit max_ipm max_admm 0 1e-10 tol 5 1 out")
        a.time_limit = float(L[0])
        a.max_ipm_iters = int(float(L[1]))
        a.max_admm_iters = int(float(L[2]))
        a.tol = float(L[5])
        a.out = L[8]
    if a.out is None:
        a.out = a.mps[:-4] if a.mps.lower().endswith(".mps") else a.mps
    return a


def main(argv=None) -> int:
    a = parse_args(sys.argv[1:] if argv is None else argv)
    from . import mps
    from .api import lp_solve
    t0 = time.perf_counter()
    std = mps.load_standard_form(a.mps)
    t_read = time.perf_counter() - t0
    params = dict(tol=a.tol, verbose=a.verbose)
    if a.time_limit is not None:
        params["timelimit"] = a.time_limit
    if a.max_ipm_iters is not None:
        params["max_ipm_iter"] = a.max_ipm_iters
    if a.max_admm_iters is not None:
        params["max_admm_iter"] = a.max_admm_iters
    t1 = time.perf_counter()
    x, y, s, info = lp_solve(std.A, std.b, std.c, params, want_stats=True)
    t_solve = time.perf_counter() - t1
    res = dict(problem=std.name or a.mps, m=int(std.A.shape[0]), n=int(std.A.shape[1]), nnz=int(std.A.nnz),
               n_original=int(std.n_orig), status=info["status"], status_val=int(info["status_val"]),
               ipm_iter=int(info["ipm_iter"]), admm_iter=int(info["admm_iter"]),
               pobj=float(info["pobj"]) + std.objcon, dobj=float(info["dobj"]) + std.objcon, objcon=std.objcon,
               pres=float(info["pres"]), dres=float(info["dres"]), gap=float(info["gap"]),
               read_time_s=t_read, solve_wall_s=t_solve, tol=a.tol, solver="abip-lp-b200")
    # the keys the reference's result scripts read (scripts/bench-lp/analyze_abip.py:10-60): MATLAB schema
    # (pres, dres, time, status, pobj, dobj, admm_iter, ipm_iter) and the schema of its C binary
    res["time"] = float(info["setup_time_ms"] + info["solve_time_ms"]) / 1e3
    res.update(InnerIter=res["admm_iter"], OuterIter=res["ipm_iter"], PResABIP=res["pres"], DResABIP=res["dres"],
               ABIPTime=res["time"], PObj=res["pobj"], DObj=res["dobj"],
               CGIter=int(info.get("stats", {}).get("n_cg_iters", 0)))
    with open(a.out + ".json", "w") as fh:
        json.dump(res, fh, indent=1)
    xo = std.recover(x)
    with open(a.out + ".sol", "w") as fh:
        fh.write("\n".join(repr(float(v)) for v in xo) + "\n")
    print(json.dumps(res))
    return 0 if info["status_val"] != -4 else 1      # ABIP_FAILED


if __name__ == "__main__":
    sys.exit(main())
