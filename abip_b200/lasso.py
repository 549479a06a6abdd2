"""Lasso front end on the ABIP-QCP engine:  min_w 1/2 |X w - y|^2 + lambda |w|_1.

The reference solves it through its "ml" entry (mex/abip_ml_mex.c:318-331, prob_type = LASSO) as a conic program with one
rotated second-order cone of dimension m + 2 and 2 n non-negative variables (source/lasso_config.c:8-92):

    variables   x = [t0, t1, z (m), w+ (n), w- (n)],   2 t0 t1 >= |z|^2,  w+, w- >= 0
    constraints t0 = 1,   z + X (w+ - w-) = y                                  (lasso_A_times, lasso_config.c:99-110)
    objective   t1 + lambda 1'(w+ + w-)                                        (= 1/2 |y - X w|^2 + lambda |w|_1)

and runs it with a problem-specific scaling, residual definition and linear-system vtable (lasso_config.c:131-720).  This
module builds the same cone program explicitly and hands it to the GENERAL QCP engine (abip_b200.qcp, K = {rq: [m + 2],
l: 2 n}): same minimiser and objective as the reference's Lasso mode, not the same iteration counts (the specialised
scaling constants and PCG tolerance schedule of lasso_config.c are not reproduced).  The returned w is x[w+] - x[w-], as in
un_scaling_lasso_sol (lasso_config.c:303-318)."""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def lasso_cone_program(X, y, lam: float):
    """(A, b, c, K) of the cone program above; X: m x n (dense or sparse), y: m."""
    X = sp.csc_matrix(X, dtype=np.float64)
    m, n = X.shape
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    if y.size != m:
        raise ValueError("y must have one entry per row of X")
    if not lam > 0:
        raise ValueError("lambda must be positive")
    top = sp.csc_matrix(([1.0], ([0], [0])), shape=(1, 2 + m + 2 * n))
    body = sp.hstack([sp.csc_matrix((m, 2)), sp.identity(m, format="csc"), X, -X], format="csc")
    A = sp.vstack([top, body], format="csc")
    A.sort_indices()
    b = np.concatenate([[1.0], y])
    c = np.concatenate([[0.0, 1.0], np.zeros(m), np.full(2 * n, float(lam))])
    K = {"rq": [m + 2], "l": 2 * n}
    return A, b, c, K


def lasso_objective(X, y, lam: float, w) -> float:
    r = X @ w - y
    return 0.5 * float(r @ r) + float(lam) * float(np.abs(w).sum())


def lasso_solve(X, y, lam: float, **settings):
    """Solve on the GPU engine.  settings: ABIP-QCP settings (eps_p, eps_d, eps_g, max_admm_iters, verbose, ...).
    Returns (w, info); info carries the engine's status / iteration counts and `objective` = the Lasso objective at w."""
    from . import qcp
    A, b, c, K = lasso_cone_program(X, y, lam)
    m, n = sp.csc_matrix(X).shape
    opts = dict(verbose=0)
    opts.update(settings)
    x, yy, s, info = qcp.qcp_solve_raw(A, None, b, c, K, **opts)
    w = x[2 + m:2 + m + n] - x[2 + m + n:]
    info["objective"] = lasso_objective(sp.csc_matrix(X), np.asarray(y, dtype=np.float64), lam, w)
    return w, info
