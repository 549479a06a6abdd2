"""Lasso front end (abip_b200/lasso.py) against the reference's own Lasso mode.

tests/golden/lasso_golden.json holds the results of the UNMODIFIED reference (oracle/_ref/libabip_qcp_ref.so, abip() with
prob_type = LASSO as driven by mex/abip_ml_mex.c:318-331; generator: tests/golden/make_golden_lasso.py).  The front end
builds the cone program of source/lasso_config.c:8-130 explicitly and solves it on the general QCP path, so the minimiser
and the objective must agree with the reference; iteration counts are compared with the oracle's general path (the
reference's Lasso mode has its own scaling and stopping constants)."""
import json
import os

import numpy as np
import pytest

from abip_b200 import lasso, problems
from oracle import qcp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "lasso_golden.json")))
EPS = 1e-5


def _oracle(X, y, lam):
    A, b, c, K = lasso.lasso_cone_program(X, y, lam)
    m, n = X.shape
    r = O.solve(A, None, b, c, K, O.Settings(eps_p=EPS, eps_d=EPS, eps_g=EPS))
    return r, r.x[2 + m:2 + m + n] - r.x[2 + m + n:]


def test_cone_program_shape_and_feasible_point():
    X, y, lam = problems.LASSO_CASES["sparse_wide"]()
    A, b, c, K = lasso.lasso_cone_program(X, y, lam)
    m, n = X.shape
    assert A.shape == (m + 1, 2 + m + 2 * n) and K == {"rq": [m + 2], "l": 2 * n}
    w = np.zeros(n)
    w[:3] = [0.5, -0.25, 1.0]
    z = y - X @ w
    x = np.concatenate([[1.0, 0.5 * z @ z], z, np.maximum(w, 0), np.maximum(-w, 0)])
    assert np.allclose(A @ x, b) and 2 * x[0] * x[1] >= z @ z - 1e-12
    assert abs(c @ x - lasso.lasso_objective(X, y, lam, w)) < 1e-12
    with pytest.raises(ValueError):
        lasso.lasso_cone_program(X, y[:-1], lam)
    with pytest.raises(ValueError):
        lasso.lasso_cone_program(X, y, 0.0)


@pytest.mark.parametrize("name", sorted(problems.LASSO_CASES))
def test_cone_program_reproduces_reference_lasso(name):
    """CPU: the oracle on the explicit cone program lands on the reference's Lasso solution."""
    X, y, lam = problems.LASSO_CASES[name]()
    g = GOLD[name]
    r, w = _oracle(X, y, lam)
    assert r.status == "Solved" == g["status"]
    assert abs(lasso.lasso_objective(X, y, lam, w) - g["objective"]) <= 1e-5 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(problems.LASSO_CASES))
def test_lasso_gpu_matches_reference(name):
    """GPU engine through lasso_solve: the reference's objective to 1e-5 and coefficients to 1e-3; status and ADMM
    iteration count (within 5 %) of the oracle's general path on the same cone program."""
    X, y, lam = problems.LASSO_CASES[name]()
    g = GOLD[name]
    w, info = lasso.lasso_solve(X, y, lam, eps_p=EPS, eps_d=EPS, eps_g=EPS)
    r, w_or = _oracle(X, y, lam)
    assert info["status"] == "Solved" == g["status"]
    assert abs(info["objective"] - g["objective"]) <= 1e-5 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))
    assert info["ipm_iter"] == r.ipm_iter
    assert abs(info["admm_iter"] - r.admm_iter) <= max(2, 0.05 * r.admm_iter)
    assert np.max(np.abs(w - w_or)) <= 1e-4 * max(1.0, np.max(np.abs(w_or)))
