"""Soft-margin SVM front end (abip_b200/svm.py) against the reference's own SVM mode.

tests/golden/svm_golden.json holds the results of the UNMODIFIED reference (oracle/_ref/libabip_qcp_ref.so, abip() with
prob_type = SVM as driven by mex/abip_ml_mex.c:332-336; generator: tests/golden/make_golden_svm.py).  The front end builds the
cone program of source/svm_config.c:8-230 explicitly and solves it on the general QCP path, so minimiser and objective must
agree with the reference (to the accuracy of its own eps = 1e-5 solve); iteration counts are compared with the oracle's
general path."""
import json
import os

import numpy as np
import pytest

from abip_b200 import problems, svm
from oracle import qcp_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "svm_golden.json")))
EPS = 1e-5


def _oracle(X, y, Cp):
    A, b, c, K = svm.svm_cone_program(X, y, Cp)
    m, n = X.shape
    r = O.solve(A, None, b, c, K, O.Settings(eps_p=EPS, eps_d=EPS, eps_g=EPS))
    return r, svm.svm_split(r.x, m, n)


def test_cone_program_shape_and_feasible_point():
    X, y, Cp = problems.SVM_CASES["dense_tall"]()
    A, b, c, K = svm.svm_cone_program(X, y, Cp)
    m, n = X.shape
    assert A.shape == (m + n + 1, 4 + 3 * n + 2 * m) and K == {"rq": [n + 2], "l": 2 + 2 * m + 2 * n}
    rng = np.random.default_rng(0)
    w, b0 = rng.standard_normal(n), 0.3
    marg = y * (X @ w + b0)
    xi, t = np.maximum(0, 1 - marg), np.maximum(0, marg - 1)
    x = np.concatenate([[1.0, 0.5 * w @ w], w, np.maximum(w, 0), [max(b0, 0)], np.maximum(-w, 0), [max(-b0, 0)], xi, t])
    assert np.allclose(A @ x, b) and abs(c @ x - svm.svm_objective(X, y, Cp, w, b0)) < 1e-12
    w2, b2, xi2 = svm.svm_split(x, m, n)
    assert np.allclose(w2, w) and abs(b2 - b0) < 1e-15 and np.allclose(xi2, xi)
    with pytest.raises(ValueError):
        svm.svm_cone_program(X, 0.5 * y, Cp)
    with pytest.raises(ValueError):
        svm.svm_cone_program(X, y, -1.0)


@pytest.mark.parametrize("name", sorted(problems.SVM_CASES))
def test_cone_program_reproduces_reference_svm(name):
    """CPU: the oracle on the explicit cone program lands on the reference's SVM solution."""
    X, y, Cp = problems.SVM_CASES[name]()
    g = GOLD[name]
    r, (w, b0, xi) = _oracle(X, y, Cp)
    assert r.status == "Solved" == g["status"]
    obj = svm.svm_objective(X, y, Cp, w, b0)
    assert abs(obj - g["objective"]) <= 1e-3 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))
    assert abs(b0 - g["b"]) <= 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(problems.SVM_CASES))
def test_svm_gpu_matches_reference(name):
    """GPU engine through svm_solve: the reference's objective to 1e-3 and (w, b) to 1e-3; status and ADMM iteration count
    (within 5 %) of the oracle's general path on the same cone program, its (w, b) to 1e-4."""
    X, y, Cp = problems.SVM_CASES[name]()
    g = GOLD[name]
    w, b0, xi, info = svm.svm_solve(X, y, Cp, eps_p=EPS, eps_d=EPS, eps_g=EPS)
    r, (w_or, b_or, _) = _oracle(X, y, Cp)
    assert info["status"] == "Solved" == g["status"]
    assert abs(info["objective"] - g["objective"]) <= 1e-3 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))
    assert info["ipm_iter"] == r.ipm_iter
    assert abs(info["admm_iter"] - r.admm_iter) <= max(2, 0.05 * r.admm_iter)
    assert np.max(np.abs(w - w_or)) <= 1e-4 * max(1.0, np.max(np.abs(w_or))) and abs(b0 - b_or) <= 1e-4
    assert np.all(xi >= -1e-9)


def _oracle_qp(X, y, lam):
    A, Q, b, c, K = svm.svm_qp_program(X, y, lam)
    m, n = X.shape
    r = O.solve(A, Q, b, c, K, O.Settings(eps_p=EPS, eps_d=EPS, eps_g=EPS))
    return r, (r.x[:n], float(r.x[n]))


@pytest.mark.parametrize("name", sorted(problems.SVM_CASES))
def test_qp_form_reproduces_reference_svmqp(name):
    """CPU: the QP form (svm_qp_config.c:8-150: (w, b) free, Q = diag(1_n, 0), hinge weight 1 / (m lambda)) on the oracle
    against the reference's SVMQP mode."""
    X, y, Cp = problems.SVM_CASES[name]()
    m = X.shape[0]
    g = GOLD[name + "_qp"]
    r, (w, b0) = _oracle_qp(X, y, 1.0 / (m * Cp))
    assert r.status == "Solved" == g["status"]
    assert abs(svm.svm_objective(X, y, Cp, w, b0) - g["objective"]) <= 1e-3 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))
    assert abs(b0 - g["b"]) <= 1e-3


def test_schur_path_handles_a_vanishing_reduced_rhs():
    """The engine's linear-system path (m-space Schur PCG, restated in the oracle) on the QP form: in the first ADMM iteration
    the reduced right-hand side b_y - A H^-1 b_x is exactly zero with a non-zero warm start; a tolerance relative to |rhs|
    alone is then 0 and CG runs into 0 / 0 (the GPU engine went on with NaN iterates until its iteration limit).  With the
    floor of 1e-13 x the warm-start residual the Schur path reproduces the direct path."""
    X, y, Cp = problems.SVM_CASES["dense_tall"]()
    m = X.shape[0]
    A, Q, b, c, K = svm.svm_qp_program(X, y, 1.0 / (m * Cp))
    d = O.solve(A, Q, b, c, K, O.Settings(eps_p=EPS, eps_d=EPS, eps_g=EPS))
    s = O.solve(A, Q, b, c, K, O.Settings(eps_p=EPS, eps_d=EPS, eps_g=EPS), linsys="schur", pcg_rtol=1e-8)
    assert (s.status, s.ipm_iter, s.admm_iter) == (d.status, d.ipm_iter, d.admm_iter)
    assert abs(s.pobj - d.pobj) <= 1e-6 * abs(d.pobj)


@pytest.mark.gpu
@pytest.mark.timeout(120)
@pytest.mark.parametrize("name", sorted(problems.SVM_CASES))
def test_svm_qp_gpu_matches_reference(name):
    """GPU engine through svm_qp_solve (general QCP path with a diagonal Q and free variables) against the reference's SVMQP
    mode and the oracle's iteration counts."""
    X, y, Cp = problems.SVM_CASES[name]()
    m = X.shape[0]
    g = GOLD[name + "_qp"]
    w, b0, xi, info = svm.svm_qp_solve(X, y, 1.0 / (m * Cp), eps_p=EPS, eps_d=EPS, eps_g=EPS, max_admm_iters=20000)
    r, (w_or, b_or) = _oracle_qp(X, y, 1.0 / (m * Cp))
    assert info["status"] == "Solved" == g["status"]
    assert abs(info["objective"] - g["objective"]) <= 1e-3 * abs(g["objective"])
    assert np.max(np.abs(w - np.array(g["w"]))) <= 1e-3 * max(1.0, np.max(np.abs(g["w"])))
    assert info["ipm_iter"] == r.ipm_iter
    assert abs(info["admm_iter"] - r.admm_iter) <= max(2, 0.05 * r.admm_iter)
    assert np.max(np.abs(w - w_or)) <= 1e-4 * max(1.0, np.max(np.abs(w_or))) and abs(b0 - b_or) <= 1e-4
