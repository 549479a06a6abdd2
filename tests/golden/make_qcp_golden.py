"""Generates tests/golden/qcp_golden.json by running the UNMODIFIED reference ABIP-QCP solver (compiled with stub
MKL headers by oracle/Makefile, linsys_solver = 1 QDLDL) on seeded synthetic QCPs and on the explicit toy problem
of the reference's own test (test/test_abip_install.m:32-43).

    make -C oracle qcp && python tests/golden/make_qcp_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from abip_b200 import problems  # noqa: E402
from oracle import ref_qcp  # noqa: E402

CASES = {
    "toy_qcp": lambda: problems.toy_qcp(),
    "mixed_cones_q": lambda: problems.random_qcp(30, 6, 5, n_rsoc=3, rsoc_dim=4, n_free=4, n_lin=10, seed=1),
    "socp_noq": lambda: problems.random_qcp(40, 10, 6, n_lin=20, seed=2, with_q=False),
    "qp_lin_only": lambda: problems.random_qcp(50, 0, 0, n_lin=150, seed=3),
    "soc_dim1_and_big": lambda: problems.random_qcp(20, 1, 60, n_lin=5, seed=4),
    "rsoc_only": lambda: problems.random_qcp(25, 0, 0, n_rsoc=12, rsoc_dim=5, seed=6),
    "cfg3_scale0.003": lambda: problems.cfg3(scale=0.003),
    "cfg3_scale0.01": lambda: problems.cfg3(scale=0.01),
}


def main():
    out = {}
    for name, gen in CASES.items():
        p = gen()
        r = ref_qcp.solve(p, eps_p=1e-4, eps_d=1e-4, eps_g=1e-4)
        out[name] = {"m": p.m, "n": p.n, "status": r["status"], "status_val": r["status_val"],
                     "ipm_iter": r["ipm_iter"], "admm_iter": r["admm_iter"], "pobj": r["pobj"], "dobj": r["dobj"],
                     "res_pri": r["res_pri"], "res_dual": r["res_dual"], "rel_gap": r["rel_gap"],
                     "x_head": r["x"][:16].tolist(), "y_head": r["y"][:16].tolist(), "s_head": r["s"][:16].tolist(),
                     "x_norm": float(np.linalg.norm(r["x"]))}
        print(name, p.m, p.n, r["status"], r["ipm_iter"], r["admm_iter"], r["pobj"])
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "qcp_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
