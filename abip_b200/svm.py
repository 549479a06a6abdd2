"""Soft-margin SVM front end on the ABIP-QCP engine:  min_{w, b, xi} 1/2 |w|^2 + C sum xi
s.t. y_i (x_i . w + b) >= 1 - xi_i, xi >= 0.

The reference solves it through its "ml" entry (mex/abip_ml_mex.c:332-336, prob_type = SVM, C = lambda) as a cone program
with one rotated second-order cone of dimension n + 2 and 2 + 2 m + 2 n non-negative variables (source/svm_config.c:8-230;
operator svm_A_times :175-196 with all scaling weights = 1):

    variables   x = [t0, t1, z (n) | w+ (n), b+, w- (n), b-, xi (m), t (m)],   2 t0 t1 >= |z|^2,  the rest >= 0
    constraints t0 = 1
                diag(y) X (w+ - w-) + y (b+ - b-) + xi - t = 1          (m rows)
                z - (w+ - w-) = 0                                       (n rows)
    objective   t1 + C 1'xi                                             (= 1/2 |w|^2 + C sum xi)

and runs it with a problem-specific scaling / residual / linear-system vtable (svm_config.c:231-1001).  This module builds
the same cone program explicitly and hands it to the GENERAL QCP engine (K = {rq: [n + 2], l: 2 + 2 m + 2 n}): same
minimiser and objective as the reference's SVM mode, not the same iteration counts.  Returned: w = w+ - w-, b = b+ - b-,
xi, as in un_scaling_svm_sol (svm_config.c:413-440)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def svm_cone_program(X, y, C: float):
    """(A, b, c, K) of the cone program above; X: m x n (samples x features), y: m labels in {-1, +1}."""
    X = sp.csc_matrix(X, dtype=np.float64)
    m, n = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    if y.size != m or not np.all(np.abs(y) == 1.0):
        raise ValueError("y must hold one label in {-1, +1} per row of X")
    if not C > 0:
        raise ValueError("C must be positive")
    YX = sp.diags(y) @ X
    ycol = sp.csc_matrix(y.reshape(-1, 1))
    q = 4 + 3 * n + 2 * m
    row0 = sp.csc_matrix(([1.0], ([0], [0])), shape=(1, q))
    Im, In = sp.identity(m, format="csc"), sp.identity(n, format="csc")
    margin = sp.hstack([sp.csc_matrix((m, 2 + n)), YX, ycol, -YX, -ycol, Im, -Im], format="csc")
    link = sp.hstack([sp.csc_matrix((n, 2)), In, -In, sp.csc_matrix((n, 1)), In, sp.csc_matrix((n, 1 + 2 * m))], format="csc")
    A = sp.vstack([row0, margin, link], format="csc")
    A.sort_indices()
    b = np.concatenate([[1.0], np.ones(m), np.zeros(n)])
    c = np.zeros(q)
    c[1] = 1.0
    c[3 * n + 4:3 * n + 4 + m] = float(C)
    K = {"rq": [n + 2], "l": 2 + 2 * m + 2 * n}
    return A, b, c, K


def svm_split(x, m: int, n: int):
    """(w, b, xi) from the solution vector of the cone program."""
    w = x[n + 2:2 * n + 2] - x[2 * n + 3:3 * n + 3]
    b = x[2 * n + 2] - x[3 * n + 3]
    xi = x[3 * n + 4:3 * n + 4 + m]
    return w, float(b), xi


def svm_objective(X, y, C: float, w, b: float) -> float:
    xi = np.maximum(0.0, 1.0 - y * (X @ w + b))
    return 0.5 * float(w @ w) + float(C) * float(xi.sum())


def svm_solve(X, y, C: float, **settings):
    """Solve on the GPU engine.  settings: ABIP-QCP settings.  Returns (w, b, xi, info); info["objective"] is the SVM
    objective (hinge losses recomputed from w, b)."""
    from . import qcp
    A, bb, c, K = svm_cone_program(X, y, C)
    Xs = sp.csc_matrix(X)
    m, n = Xs.shape
    opts = dict(verbose=0)
    opts.update(settings)
    x, yy, s, info = qcp.qcp_solve_raw(A, None, bb, c, K, **opts)
    w, b0, xi = svm_split(x, m, n)
    info["objective"] = svm_objective(Xs, np.asarray(y, dtype=np.float64), C, w, b0)
    return w, b0, xi, info


# ---------------------------------------------------------------------------------------------------------
# QP form (prob_type = SVMQP, mex/abip_ml_mex.c:337-342; source/svm_qp_config.c:8-150):
#     min 1/2 |w|^2 + 1/(m lambda) sum xi   s.t.   diag(y) X w + y b + xi - t = 1,  (w, b) free,  xi, t >= 0
# variables x = [w (n), b | xi (m), t (m)], K = {f: n + 1, l: 2 m}, Q = diag(1_n, 0): the quadratic term goes to Q instead of
# a rotated cone.  (In the first ADMM iteration of this program the reduced right-hand side of the Schur system is exactly
# zero: the PCG tolerance of the engine is floored by 1e-13 x the warm-start residual for that case, qcp_engine.cu.)
# ---------------------------------------------------------------------------------------------------------
def svm_qp_program(X, y, lam: float):
    """(A, Q, b, c, K) of the QP form; the weight of the hinge losses is 1 / (m lambda) as in the reference."""
    X = sp.csc_matrix(X, dtype=np.float64)
    m, n = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    if y.size != m or not np.all(np.abs(y) == 1.0):
        raise ValueError("y must hold one label in {-1, +1} per row of X")
    if not lam > 0:
        raise ValueError("lambda must be positive")
    Im = sp.identity(m, format="csc")
    A = sp.hstack([sp.diags(y) @ X, sp.csc_matrix(y.reshape(-1, 1)), Im, -Im], format="csc")
    A.sort_indices()
    q = 1 + n + 2 * m
    Q = sp.diags(np.concatenate([np.ones(n), np.zeros(q - n)]), format="csc")
    c = np.zeros(q)
    c[n + 1:n + 1 + m] = 1.0 / (m * float(lam))
    K = {"f": n + 1, "l": 2 * m}
    return A, Q, np.ones(m), c, K


def svm_qp_solve(X, y, lam: float, **settings):
    """QP form on the GPU engine; returns (w, b, xi, info) with info["objective"] = 1/2 |w|^2 + sum(hinge) / (m lambda)."""
    from . import qcp
    A, Q, bb, c, K = svm_qp_program(X, y, lam)
    Xs = sp.csc_matrix(X)
    m, n = Xs.shape
    opts = dict(verbose=0, max_admm_iters=200000)
    opts.update(settings)
    x, yy, s, info = qcp.qcp_solve_raw(A, Q, bb, c, K, **opts)
    w, b0, xi = x[:n], float(x[n]), x[n + 1:n + 1 + m]
    info["objective"] = svm_objective(Xs, np.asarray(y, dtype=np.float64), 1.0 / (m * float(lam)), w, b0)
    return w, b0, xi, info
