"""Generates tests/golden/lp_golden.json by running the UNMODIFIED reference ABIP-LP solver, compiled by
oracle/Makefile into oracle/_ref/libabip_indirect_ref.so, on the seeded synthetic problems of
abip_b200/problems.py.  Run in the build container (needs /root/reference for the oracle build):

    make -C oracle lp && python tests/golden/make_golden.py

The fixtures pin (a) the numpy oracle (tests/test_oracle_vs_ref.py) and (b) the GPU engine
(tests/test_lp_gpu.py) to the reference's own outputs; nothing under /root/reference is read at test time.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from abip_b200 import problems  # noqa: E402
from oracle import ref_lp  # noqa: E402

CASES = {
    "cfg1": (lambda: problems.cfg1(), dict(eps=1e-4)),
    "rand_200x700": (lambda: problems.random_lp(200, 700, 4, seed=3), dict(eps=1e-4)),
    "rand_200x700_eps1e-3": (lambda: problems.random_lp(200, 700, 4, seed=3), dict(eps=1e-3)),
    "rand_200x700_half": (lambda: problems.random_lp(200, 700, 4, seed=3), dict(eps=1e-4, half_update=1)),
    "rand_200x700_noadapt": (lambda: problems.random_lp(200, 700, 4, seed=3), dict(eps=1e-4, adaptive=0)),
    "rand_200x700_nonorm": (lambda: problems.random_lp(200, 700, 4, seed=3), dict(eps=1e-4, normalize=0)),
    "rand_1x9": (lambda: problems.random_lp(1, 9, 1, seed=4), dict(eps=1e-3)),
    "rand_37x1000_dense_rows": (lambda: problems.random_lp(37, 1000, 20, seed=5), dict(eps=1e-4)),
    "mcf_small": (lambda: problems.mcf_lp(4, 40, 200, 6, 300, seed=6), dict(eps=1e-4)),
    "cfg2_scale0.01": (lambda: problems.cfg2(scale=0.01), dict(eps=1e-4)),
    "cfg5_lp_0": (lambda: problems.random_lp(500, 2000, 5, seed=5000), dict(eps=1e-4)),
}


def main():
    out = {}
    for name, (gen, kw) in CASES.items():
        p = gen()
        r = ref_lp.solve(p, **kw)
        out[name] = {
            "m": p.m, "n": p.n, "nnz": p.nnz, "settings": kw,
            "status": r["status"], "status_val": r["status_val"], "ipm_iter": r["ipm_iter"],
            "admm_iter": r["admm_iter"], "pobj": r["pobj"], "dobj": r["dobj"], "res_pri": r["res_pri"],
            "res_dual": r["res_dual"], "rel_gap": r["rel_gap"],
            "x_head": r["x"][:16].tolist(), "y_head": r["y"][:16].tolist(), "s_head": r["s"][:16].tolist(),
            "x_norm": float(np.linalg.norm(r["x"])), "y_norm": float(np.linalg.norm(r["y"])),
            "s_norm": float(np.linalg.norm(r["s"])),
        }
        print(name, out[name]["status"], out[name]["ipm_iter"], out[name]["admm_iter"], out[name]["pobj"])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lp_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
