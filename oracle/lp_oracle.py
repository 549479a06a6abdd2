"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference ABIP-LP indirect path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (abip_b200/) never does.  Parity status: the reference's own tests hold no golden vectors for
this path (test/test_abip_install.m has no assertions), so this restatement is pinned
against outputs of the reference itself, compiled unmodified by oracle/Makefile into oracle/_ref
(tests/test_oracle_vs_ref.py) and against the fixtures in tests/golden/ generated from it
(tests/golden/make_golden.py).

Every function cites the reference file:line it restates (paths relative to
/root/reference/src/abip-lp).  Summation order differs from the serial C loops (numpy pairwise sums),
so agreement is to rounding, not bit-exact.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import numpy as np
import scipy.sparse as sp

MIN_SCALE = 1e-3   # linsys/common.c:4-5
MAX_SCALE = 1e3
CG_BEST_TOL = 1e-9  # linsys/indirect.c:3-4
CG_MIN_TOL = 1e-1
EPS_TOL = 1e-18     # include/glbopts.h:157-158
INDETERMINATE_TOL = 1e-9

ABIP_SOLVED = 1
ABIP_SOLVED_INACCURATE = 2
ABIP_UNBOUNDED = -1
ABIP_INFEASIBLE = -2
ABIP_INDETERMINATE = -3
ABIP_UNFINISHED = 0


def safediv_pos(x, y):
    """glbopts.h:158 SAFEDIV_POS."""
    return x / EPS_TOL if y < EPS_TOL else x / y


@dataclass
class Settings:
    """src/util.c:288-329 set_default_settings + mex-only defaults (mexfile/abip_mex.c:320-341)."""
    normalize: int = 1
    pfeasopt: int = 0
    scale: float = 1.0
    rho_y: float = 1e-3
    sparsity_ratio: float = 0.01
    max_ipm_iters: int = 500
    max_admm_iters: int = 1000000
    max_time: float = 3600.0
    eps: float = 1e-3
    alpha: float = 1.8
    cg_rate: float = 2.0
    adaptive: int = 1
    eps_cor: float = 0.2
    eps_pen: float = 0.1
    dynamic_sigma: float = -1.0
    dynamic_x: float = 0.8
    dynamic_eta: float = 1.1
    restart_fre: int = 1000
    restart_thresh: int = 100000
    verbose: int = 0
    warm_start: int = 0
    adaptive_lookback: int = 20
    origin_rescale: int = 0
    pc_ruiz_rescale: int = 1
    qp_rescale: int = 0
    ruiz_iter: int = 10
    hybrid_mu: int = 1
    hybrid_thresh: float = 1000.0
    dynamic_sigma_second: float = 0.5
    half_update: int = 0
    avg_criterion: int = 0


# --------------------------------------------------------------------------------------------------
# linsys/common.c
# --------------------------------------------------------------------------------------------------
def normalize_A(A: sp.csc_matrix, stgs: Settings):
    """linsys/common.c:150-565 _normalize_A (pc + origin + ruiz + qp).  Returns scaled A, D, E,
    mean_norm_row_A, mean_norm_col_A."""
    A = A.tocsc(copy=True).astype(np.float64)
    m, n = A.shape
    Ap, Ai = A.indptr, A.indices
    colidx = np.repeat(np.arange(n), np.diff(Ap))
    min_row, max_row = MIN_SCALE * math.sqrt(n), MAX_SCALE * math.sqrt(n)
    min_col, max_col = MIN_SCALE * math.sqrt(m), MAX_SCALE * math.sqrt(m)

    def clampc(e):
        e = e.copy()
        e[e < min_col] = 1.0
        return np.minimum(e, max_col)

    def clampr(d):
        d = d.copy()
        d[d < min_row] = 1.0
        return np.minimum(d, max_row)

    def colreduce(v, op):
        out = np.zeros(n)
        nz = np.diff(Ap) > 0
        out[nz] = op.reduceat(v, Ap[:-1][nz])
        return out

    D_pc, E_pc = np.ones(m), np.ones(n)
    D_or, E_or = np.ones(m), np.ones(n)
    D_ru, E_ru = np.ones(m), np.ones(n)
    D_qp, E_qp = np.ones(m), np.ones(n)
    x = A.data
    if stgs.pc_ruiz_rescale:  # :216-268
        E_pc = clampc(np.sqrt(colreduce(np.abs(x), np.add)))
        x /= E_pc[colidx]
        D_pc = clampr(np.sqrt(np.bincount(Ai, weights=np.abs(x), minlength=m)))
        x /= D_pc[Ai]
    if stgs.origin_rescale:  # :282-332
        E_or = clampc(np.sqrt(colreduce(x * x, np.add)))
        x /= E_or[colidx]
        D_or = clampr(np.sqrt(np.bincount(Ai, weights=x * x, minlength=m)))
        x /= D_or[Ai]
    if stgs.pc_ruiz_rescale:  # :345-417
        for _ in range(stgs.ruiz_iter):
            Et = clampc(np.sqrt(colreduce(np.abs(x), np.maximum)))
            x /= Et[colidx]
            Dt = np.zeros(m)
            np.maximum.at(Dt, Ai, np.abs(x))
            Dt = clampr(np.sqrt(Dt))
            x /= Dt[Ai]
            E_ru *= Et
            D_ru *= Dt
    if stgs.qp_rescale:  # :419-512
        mx = colreduce(np.abs(x), np.maximum)
        ax = np.abs(x)
        axp = np.where(ax > 0, ax, np.inf)
        mn = np.minimum(colreduce(axp, np.minimum), mx)
        E_qp = clampc(np.sqrt(mn) * np.sqrt(mx))
        x /= E_qp[colidx]
        ax = np.abs(x)
        Dmax = np.zeros(m)
        np.maximum.at(Dmax, Ai, ax)
        Dmin = Dmax.copy()
        np.minimum.at(Dmin, Ai, np.where(ax > 0, ax, np.inf))
        D_qp = clampr(np.sqrt(Dmax * Dmin))
        x /= D_qp[Ai]
    D = D_pc * D_ru * D_or * D_qp  # :524-532
    E = E_pc * E_ru * E_or * E_qp
    nms = np.bincount(Ai, weights=x * x, minlength=m)
    mean_row = float(np.sum(np.sqrt(nms) / m))  # :535-548
    mean_col = float(np.sum(np.sqrt(colreduce(x * x, np.add)) / n))  # :552-557
    if stgs.scale != 1:
        x *= stgs.scale
    return A, D, E, mean_row, mean_col


def accum_by_Atrans(A: sp.csc_matrix, x, y):
    """linsys/common.c:598-639: y += A'x (gather over CSC columns)."""
    y += A.T @ x


def accum_by_A(At_csr: sp.csr_matrix, x, y):
    """linsys/indirect.c:233-242: y += A x, executed as the gather kernel on the stored transpose."""
    y += At_csr @ x


# --------------------------------------------------------------------------------------------------
# linsys/indirect.c
# --------------------------------------------------------------------------------------------------
class LinSys:
    """struct ABIP_LIN_SYS_WORK (linsys/indirect.h:14-29) + init (indirect.c:282-318)."""

    def __init__(self, A: sp.csc_matrix, stgs: Settings):
        self.A = A.tocsc()
        self.Acsr = A.tocsr()  # transpose(): indirect.c:81-139 (CSC of A' == CSR of A)
        self.m, self.n = A.shape
        self.stgs = stgs
        # get_preconditioner indirect.c:36-79: M = 1/diag(AA') -- note: no rho_y term
        self.M = 1.0 / np.asarray(self.A.multiply(self.A).sum(axis=1)).ravel()
        self.tot_cg_its = 0
        self.n_solves = 0
        self.n_matvec = 0

    def mat_vec(self, x):
        """indirect.c:205-220: y = A(A'x) + rho_y x."""
        self.n_matvec += 1
        return self.Acsr @ (self.A.T @ x) + self.stgs.rho_y * x

    def pcg(self, s, b, max_its, tol):
        """indirect.c:321-391.  Returns (solution, iterations)."""
        m = self.m
        if s is None:
            r = b.copy()
            x = np.zeros(m)
        else:
            r = b - self.mat_vec(s[:m])
            x = s[:m].copy()
        if np.linalg.norm(r) < min(tol, 1e-18):
            return x, 0
        z = r * self.M
        ipzr = float(z @ r)
        p = z.copy()
        i = 0
        for i in range(max_its):
            Gp = self.mat_vec(p)
            alpha = ipzr / float(p @ Gp)
            x += alpha * p
            r -= alpha * Gp
            if np.linalg.norm(r) < tol:
                return x, i + 1
            ipzr_old = ipzr
            z = r * self.M
            ipzr = float(z @ r)
            p = p * (ipzr / ipzr_old) + z
        return x, max_its

    def solve(self, b, s, it):
        """indirect.c:393-434 solve_lin_sys: b (length m+n) is overwritten with the solution."""
        m, n = self.m, self.n
        cg_tol = np.linalg.norm(b[:m]) * (CG_BEST_TOL if it < 0 else CG_MIN_TOL / (it + 1.0) ** self.stgs.cg_rate)
        cg_tol = max(cg_tol, 1e-7)
        b[:m] += self.Acsr @ b[m:m + n]
        y, its = self.pcg(s, b[:m], m, max(cg_tol, CG_BEST_TOL))
        b[:m] = y
        b[m:m + n] = -b[m:m + n] + self.A.T @ y
        if it >= 0:
            self.tot_cg_its += its
        self.n_solves += 1
        return its


# --------------------------------------------------------------------------------------------------
# src/abip.c -- work struct and per-iteration functions
# --------------------------------------------------------------------------------------------------
@dataclass
class Residuals:
    """struct ABIP_RESIDUALS include/abip.h:178-196."""
    last_ipm_iter: int = -1
    last_admm_iter: int = -1
    res_pri: float = float("nan")
    res_dual: float = float("nan")
    rel_gap: float = float("nan")
    res_infeas: float = float("nan")
    res_unbdd: float = float("nan")
    ct_x_by_tau: float = float("nan")
    bt_y_by_tau: float = float("nan")
    tau: float = float("nan")
    kap: float = float("nan")


class Work:
    """struct ABIP_WORK (include/abip.h:126-176), init_work (src/abip.c:1739-1841), update_work (:1843-1927)."""

    def __init__(self, A: sp.csc_matrix, b, c, stgs: Settings, sp_ratio=None):
        self.stgs = stgs
        self.m, self.n = A.shape
        m, n = self.m, self.n
        self.l = l = m + n + 1
        self.sp = sp_ratio if sp_ratio is not None else A.nnz / (float(m) * float(n))
        if stgs.normalize:
            self.A, self.D, self.E, self.mean_row, self.mean_col = normalize_A(A, stgs)
        else:
            self.A, self.D, self.E = A.tocsc(copy=True), None, None
        self.p = LinSys(self.A, stgs)
        z = lambda k: np.zeros(k)
        self.u, self.v, self.u_t, self.u_prev, self.v_prev = z(l), z(l), z(l), z(l), z(l)
        self.u_avg, self.v_avg, self.u_avgcon, self.v_avgcon = z(l), z(l), z(l), z(l)
        self.u_sumcon, self.v_sumcon = z(l), z(l)
        self.fre_old = 0
        # update_work
        self.nm_b = float(np.linalg.norm(b))
        self.nm_c = float(np.linalg.norm(c))
        self.b = np.array(b, dtype=np.float64)
        self.c = np.array(c, dtype=np.float64)
        self.sc_b = self.sc_c = 1.0
        if stgs.normalize:
            self.normalize_b_c()
        spmin, spmax = min(self.sp, stgs.sparsity_ratio), max(self.sp, stgs.sparsity_ratio)
        if spmax > 0.4 or (0.1 < spmin < 0.2):  # :1886-1900
            self.sigma, self.gamma = 0.3, 2.0
        elif spmin > 0.2:
            self.sigma, self.gamma = 0.5, 3.0
        else:
            self.sigma, self.gamma = 0.8, 3.0
        self.final_check = 0
        self.double_check = 0
        self.mu = 1.0
        self.beta = 1.0
        # cold_start_vars :361-381
        self.u[m:] = math.sqrt(self.mu / self.beta)
        self.v[m:] = math.sqrt(self.mu / self.beta)
        self.h = np.concatenate([-self.b, self.c])  # :1917-1919
        self.g = self.h.copy()
        self.p.solve(self.g, None, -1)  # :1922
        self.g[m:] *= -1.0
        self.g_th = float(self.h @ self.g)

    def normalize_b_c(self):
        """src/normalize.c:11-40."""
        s = self.stgs
        self.c = self.c / self.E
        nm = np.linalg.norm(self.c)
        self.sc_c = self.mean_row / max(nm, MIN_SCALE)
        self.b = self.b / self.D
        nm = np.linalg.norm(self.b)
        self.sc_b = self.mean_col / max(nm, MIN_SCALE)
        self.c = self.c * (self.sc_c * s.scale)
        self.b = self.b * (self.sc_b * s.scale)


def lin_sys_rhs(w: Work, u, v):
    """First half of project_lin_sys (src/abip.c:551-558): builds the right-hand side in place of u_t."""
    m, n, l = w.m, w.n, w.l
    ut = u + v
    ut[:m] *= w.stgs.rho_y
    ut[:l - 1] -= ut[l - 1] * w.h
    ut[:l - 1] -= w.h * (float(ut[:l - 1] @ w.g) / (w.g_th + 1))
    ut[m:l - 1] *= -1.0
    return ut


def project_lin_sys(w: Work, it: int):
    """src/abip.c:539-562."""
    l = w.l
    w.u_t = lin_sys_rhs(w, w.u, w.v)
    its = w.p.solve(w.u_t, w.u, it)
    w.u_t[l - 1] += float(w.u_t[:l - 1] @ w.h)
    return its


def barrier_prox(t, lam):
    """src/abip.c:742-746: t/2 + sqrt(t^2/4 + mu/beta)."""
    h = t / 2
    return h + np.sqrt(h * h + lam)


def project_barrier(w: Work):
    """src/abip.c:717-748."""
    m, a = w.m, w.stgs.alpha
    w.u[:m] = w.u_t[:m] - w.v[:m]
    t = a * w.u_t[m:] + (1 - a) * w.u_prev[m:] - w.v[m:]
    w.u[m:] = barrier_prox(t, w.mu / w.beta)


def update_dual_vars(w: Work):
    """src/abip.c:567-584."""
    m, a = w.m, w.stgs.alpha
    w.v[m:] += w.u[m:] - a * w.u_t[m:] - (1.0 - a) * w.u_prev[m:]


def half_update_dual_vars(w: Work):
    """src/abip.c:663-679."""
    w.v += 0.5 * (w.u - w.u_t)


def project_barrier_dual(w: Work):
    """src/abip.c:681-711."""
    m = w.m
    w.u[:] = w.u_t - w.v
    w.u[m:] = barrier_prox(w.u[m:], w.mu / w.beta)
    w.v += w.u - w.u_t


def restart_vars(w: Work, admm_iter: int, total_admm_iter: int):
    """src/abip.c:587-630."""
    fre = w.stgs.restart_fre
    w.u_avg += w.u
    w.v_avg += w.v
    if total_admm_iter < w.stgs.restart_thresh or (admm_iter + 1 - w.fre_old) % fre != 0:
        return
    w.u[:] = w.u_avg / fre
    w.v[:] = w.v_avg / fre
    w.u_avg[:] = 0
    w.v_avg[:] = 0
    w.fre_old = fre


def compute_avg(w: Work, admm_iter: int):
    """src/abip.c:635-659."""
    dom = admm_iter + 1
    w.u_sumcon += w.u
    w.v_sumcon += w.v
    w.u_avgcon[:] = w.u_sumcon / dom
    w.v_avgcon[:] = w.v_sumcon / dom


def q_norm_parts(w: Work, u, v):
    """The pieces of src/abip.c:1964-1992 for one (u, v) pair: returns (Qres_squared, 1 + sqrt(|u|^2+|v|^2))."""
    m, n = w.m, w.n
    y, x, s = u[:m], u[m:m + n], v[m:m + n]
    tau, kap = u[m + n], v[m + n]
    pr = w.p.Acsr @ x
    dr = w.A.T @ y + s
    q = float(np.sum((pr - w.b * tau) ** 2) + np.sum((dr - w.c * tau) ** 2))
    cTx = float(x @ w.c)
    bTy = float(y @ w.b)
    q += (bTy - cTx - kap) ** 2
    norm = 1 + math.sqrt(float(u @ u) + float(v @ v))
    return q, norm


def iterate_Q_norm_resd(w: Work, j: int):
    """src/abip.c:1951-2051.  Sets stgs.avg_criterion as a side effect (parity trap 5)."""
    Qres, norm = q_norm_parts(w, w.u, w.v)
    Qres_avg, norm_avg = float(w.stgs.max_admm_iters), 1.0
    if (j + 1) % 10 == 0:
        Qres_avg, norm_avg = q_norm_parts(w, w.u_avgcon, w.v_avgcon)
    if math.sqrt(Qres_avg) / norm_avg < math.sqrt(Qres) / norm:
        w.stgs.avg_criterion = 1
        return math.sqrt(Qres_avg) / norm_avg
    w.stgs.avg_criterion = 0
    return math.sqrt(Qres) / norm


def calc_residuals(w: Work, r: Residuals, ipm_iter: int, admm_iter: int):
    """src/abip.c:385-535 (calc_primal_resid, calc_dual_resid, calc_residuals)."""
    m, n, s_ = w.m, w.n, w.stgs
    if s_.avg_criterion:
        uu, vv = w.u_avgcon, w.v_avgcon
    else:
        uu, vv = w.u, w.v
    y, x, s = uu[:m], uu[m:m + n], vv[m:m + n]
    if admm_iter and r.last_admm_iter == admm_iter:
        return
    r.last_ipm_iter, r.last_admm_iter = ipm_iter, admm_iter
    nrm = (s_.scale * w.sc_c * w.sc_b) if s_.normalize else 1.0
    r.tau = abs(uu[n + m])
    r.kap = abs(vv[n + m]) / nrm
    pr = w.p.Acsr @ x
    sc = (w.D / (w.sc_b * s_.scale)) ** 2 if s_.normalize else np.ones(m)
    nm_A_x = math.sqrt(float(np.sum(pr * pr * sc)))
    nmpr = math.sqrt(float(np.sum((pr - w.b * r.tau) ** 2 * sc)))
    dr = w.A.T @ y + s
    sc = (w.E / (w.sc_c * s_.scale)) ** 2 if s_.normalize else np.ones(n)
    nm_At_ys = math.sqrt(float(np.sum(dr * dr * sc)))
    nmdr = math.sqrt(float(np.sum((dr - w.c * r.tau) ** 2 * sc)))
    r.bt_y_by_tau = float(y @ w.b) / nrm
    r.ct_x_by_tau = float(x @ w.c) / nrm
    r.res_infeas = w.nm_b * nm_At_ys / r.bt_y_by_tau if r.bt_y_by_tau > 0 else float("nan")
    r.res_unbdd = w.nm_c * nm_A_x / -r.ct_x_by_tau if r.ct_x_by_tau < 0 else float("nan")
    bt_y = safediv_pos(r.bt_y_by_tau, r.tau)
    ct_x = safediv_pos(r.ct_x_by_tau, r.tau)
    r.res_pri = safediv_pos(nmpr / (1 + w.nm_b), r.tau)
    r.res_dual = safediv_pos(nmdr / (1 + w.nm_c), r.tau)
    r.rel_gap = abs(ct_x - bt_y) / (1 + abs(ct_x) + abs(bt_y))


def has_converged(w: Work, r: Residuals, ipm_iter: int, admm_iter: int) -> int:
    """src/abip.c:1613-1641."""
    eps = w.stgs.eps
    if r.res_pri < eps and (r.res_dual < eps or w.stgs.pfeasopt) and r.rel_gap < eps:
        return ABIP_SOLVED
    if r.res_unbdd < eps and ipm_iter > 0 and admm_iter > 0:
        return ABIP_UNBOUNDED
    if r.res_infeas < eps and ipm_iter > 0 and admm_iter > 0:
        return ABIP_INFEASIBLE
    return 0


def update_barrier(w: Work, r: Residuals):
    """src/abip.c:753-921 (table-driven mu strategy)."""
    s_ = w.stgs
    ratio = w.mu / s_.eps
    err_ratio = max(max(r.res_pri, r.res_dual), r.rel_gap) / s_.eps
    dense = max(w.sp, s_.sparsity_ratio) > 0.4 or min(w.sp, s_.sparsity_ratio) > 0.1

    def gam(first):
        for lo, g in ((10.0, first), (1.0, 1.0), (0.5, 0.9), (0.1, 0.8), (0.05, 0.7), (0.01, 0.6),
                      (0.005, 0.5), (0.001, 0.4)):
            if ratio > lo:
                return g
        return 0.3

    if dense:
        gamma = gam(2.0)
        if 6 < err_ratio <= 10:
            sigma = 0.5
        elif 3 < err_ratio <= 6:
            sigma, gamma = 0.6, gamma * 0.8
        elif 1 < err_ratio <= 3:
            w.final_check = 1
            gamma *= 0.4
            sigma = 0.8 if ratio < 0.1 else 0.7
        else:
            sigma = w.sigma
    else:
        gamma = gam(3.0)
        if 6 < err_ratio <= 10:
            sigma, gamma = 0.82, gamma * 0.8
        elif 4 < err_ratio <= 6:
            sigma, gamma = 0.84, gamma * 0.6
        elif 3 < err_ratio <= 4:
            sigma, gamma = 0.85, gamma * 0.5
            w.final_check = 1
        elif 1 < err_ratio <= 3:
            w.final_check = 1
            if ratio < 0.1:
                if w.double_check:
                    sigma, gamma, w.double_check = 0.9, gamma * 0.4, 0
                else:
                    sigma, gamma, w.double_check = 1.0, gamma * 0.1, 1
            else:
                sigma, gamma = 0.88, gamma * 0.4
        else:
            sigma = w.sigma
    w.mu *= sigma
    w.sigma, w.gamma = sigma, gamma


def update_barrier_dynamic(w: Work, r: Residuals):
    """src/abip.c:930-977 (LOQO rule)."""
    m, n = w.m, w.n
    u, v = (w.u_avgcon, w.v_avgcon) if w.stgs.avg_criterion else (w.u, w.v)
    xs = u[m:] * v[m:]
    minxs = min(float(xs.min()), 1e10)
    assert minxs > 0.0, "Invalid xisi < 0"
    ksi = minxs / (float(xs.sum()) / (n + 1))
    sigma = min(0.05 * (1 - ksi) / ksi, 2.0)
    sigma = max(0.1 * sigma ** 3, w.stgs.dynamic_sigma)
    w.mu *= sigma


def update_barrier_dynamic_2(w: Work):
    """src/abip.c:982-992; note eta = dynamic_sigma (parity trap 6)."""
    w.mu *= min(w.stgs.dynamic_x * w.mu, w.mu ** w.stgs.dynamic_sigma)


def update_mu(w: Work, r: Residuals):
    """Selection logic src/abip.c:2251-2277."""
    s_ = w.stgs
    if s_.hybrid_mu:
        if s_.dynamic_sigma_second > 0.0 and w.mu < s_.hybrid_thresh * s_.eps:
            s_.dynamic_sigma = s_.dynamic_sigma_second
            update_barrier_dynamic(w, r)
        elif s_.dynamic_sigma_second == 0.0 and w.mu < s_.hybrid_thresh * s_.eps:
            s_.dynamic_sigma = s_.dynamic_sigma_second
            update_barrier(w, r)
        elif s_.dynamic_sigma < 0.0:
            update_barrier_dynamic_2(w)
    else:
        if s_.dynamic_sigma == 0.0:
            update_barrier(w, r)
        elif s_.dynamic_sigma < 0.0:
            update_barrier_dynamic_2(w)
        else:
            update_barrier_dynamic(w, r)


def reinitialize_vars(w: Work, indx: int):
    """src/abip.c:996-1075."""
    m = w.m
    u, v = (w.u_avgcon, w.v_avgcon) if w.stgs.avg_criterion else (w.u, w.v)
    if indx == 0:
        big = u[m:] > v[m:]
        v[m:][big] *= w.sigma
        u[m:][~big] *= w.sigma
    elif indx == 1:
        u[m:] *= math.sqrt(w.sigma)
        v[m:] *= math.sqrt(w.sigma)
    else:
        u[m:] *= math.sqrt(1.0 / w.sigma)
        v[m:] *= math.sqrt(1.0 / w.sigma)


# --------------------------------------------------------------------------------------------------
# src/adaptive.c
# --------------------------------------------------------------------------------------------------
def bb_half_step(w: Work, u_prev, v_prev, beta_prev, it):
    """One ADMM step as written inside update_adapt_params (src/adaptive.c:89-123): returns ut, u, v."""
    m, l, a = w.m, w.l, w.stgs.alpha
    ut = lin_sys_rhs(w, u_prev, v_prev)
    w.p.solve(ut, u_prev, it)
    ut[l - 1] += float(ut[:l - 1] @ w.h)
    u = np.empty(l)
    u[:m] = ut[:m] - v_prev[:m]
    t = a * ut[m:] + (1 - a) * u_prev[m:] - v_prev[m:]
    u[m:] = barrier_prox(t, w.mu / beta_prev)
    return ut, u, t


def bb_coefficients(stgs: Settings, delta_ut, delta_u, delta_v, beta_prev):
    """src/adaptive.c:170-223: spectral step from the 5 dots + 3 norms.  Returns the candidate beta."""
    utut = float(delta_ut @ delta_ut)
    utv = float(delta_ut @ delta_v)
    uu = float(delta_u @ delta_u)
    vv = float(delta_v @ delta_v)
    uv = float(delta_u @ delta_v)
    return bb_beta_from_scalars(stgs, utut, utv, uu, vv, uv, beta_prev)


def bb_beta_from_scalars(stgs: Settings, utut, utv, uu, vv, uv, beta_prev):
    norm_ut, norm_u, norm_v = math.sqrt(utut), math.sqrt(uu), math.sqrt(vv)
    with np.errstate(all="ignore"):
        alpha_SD = np.float64(vv) / utv
        alpha_MG = np.float64(utv) / utut
        gamma_SD = np.float64(vv) / uv
        gamma_MG = np.float64(uv) / uu
        alpha_ss = alpha_MG if 2 * alpha_MG > alpha_SD else alpha_SD - 0.5 * alpha_MG
        gamma_ss = gamma_MG if 2 * gamma_MG > gamma_SD else gamma_SD - 0.5 * gamma_MG
        alpha_cor = np.float64(utv) / (norm_v * norm_ut)
        gamma_cor = np.float64(uv) / (norm_v * norm_u)
        if alpha_cor > stgs.eps_cor and gamma_cor > stgs.eps_cor:
            beta = math.sqrt(alpha_ss * gamma_ss)
        elif alpha_cor > stgs.eps_cor and gamma_cor <= stgs.eps_cor:
            beta = float(alpha_ss)
        elif alpha_cor <= stgs.eps_cor and gamma_cor > stgs.eps_cor:
            beta = float(gamma_ss)
        else:
            beta = beta_prev
    return beta


def update_adapt_params(w: Work, it: int, trace=None):
    """src/adaptive.c:34-256 (Barzilai-Borwein search for beta)."""
    m, l, a = w.m, w.l, w.stgs.alpha
    u_prev, v_prev = w.u.copy(), w.v.copy()
    beta_prev, beta = 1.0, 0.0
    v = np.zeros(l)       # calloc'ed once (adaptive.c:282); entries [0,m) are never written
    v_next = np.zeros(l)
    for _ in range(w.stgs.adaptive_lookback):
        ut, u, t = bb_half_step(w, u_prev, v_prev, beta_prev, it)
        v[m:] = v_prev[m:] + (u[m:] - t - v_prev[m:])  # :120-123  (t = a*ut+(1-a)*u_prev - v_prev)
        ut_next, u_next, t2 = bb_half_step(w, u, v, beta_prev, it)
        v_next[m:] = v[m:] + (u_next[m:] - t2 - v[m:])  # :153-156
        delta_ut = 2.0 * v + u_next - u - v_next - v_prev  # :158-163
        delta_u = u - u_next  # :165-166
        delta_v = (u_next - u) * (a - 1.0) + v_next - v  # :168-172
        beta = bb_coefficients(w.stgs, delta_ut, delta_u, delta_v, beta_prev)
        if trace is not None:
            trace.append(beta)
        d = abs(beta - beta_prev)
        if 0 < d <= w.stgs.eps_pen:  # :225-229
            beta = (beta + beta_prev) / 2
            break
        elif d > w.stgs.eps_pen:  # :230-242
            beta_prev = beta
            u_prev = u.copy()
            v_prev = v_prev.copy()
            v_prev[:m] = v[:m]
            v_prev[m:] = (w.mu / beta_prev) / u_prev[m:]
        else:  # :243-247
            u_prev = u.copy()
            v_prev = v.copy()
    w.beta = beta


# --------------------------------------------------------------------------------------------------
# ABIP(solve) src/abip.c:2056-2297
# --------------------------------------------------------------------------------------------------
@dataclass
class Result:
    status_val: int = 0
    status: str = ""
    ipm_iter: int = 0
    admm_iter: int = 0
    pobj: float = float("nan")
    dobj: float = float("nan")
    res_pri: float = float("nan")
    res_dual: float = float("nan")
    rel_gap: float = float("nan")
    x: np.ndarray | None = None
    y: np.ndarray | None = None
    s: np.ndarray | None = None
    tot_cg_its: int = 0
    n_solves: int = 0
    trace: list = field(default_factory=list)


def get_solution(w: Work, r: Residuals, status_val: int, i: int, k: int) -> Result:
    """src/abip.c:1344-1414 get_solution + solved/infeasible/unbounded + un_normalize_sol (normalize.c:133-158)."""
    m, n, l, s_ = w.m, w.n, w.l, w.stgs
    calc_residuals(w, r, i, k)
    uu, vv = (w.u_avgcon, w.v_avgcon) if s_.avg_criterion else (w.u, w.v)
    x, y, s = uu[m:m + n].copy(), uu[:m].copy(), vv[m:m + n].copy()
    res = Result()
    nan = float("nan")
    if status_val == ABIP_UNFINISHED:
        if r.tau > INDETERMINATE_TOL and r.tau > r.kap:
            kind = "solved"
        elif np.linalg.norm(uu) < INDETERMINATE_TOL * math.sqrt(l):
            kind = "indeterminate"
        elif -r.bt_y_by_tau < r.ct_x_by_tau:
            kind = "infeasible"
        else:
            kind = "unbounded"
    elif status_val in (ABIP_SOLVED, ABIP_SOLVED_INACCURATE):
        kind = "solved"
    elif status_val in (ABIP_INFEASIBLE, -7):
        kind = "infeasible"
    else:
        kind = "unbounded"
    inacc = status_val == 0
    if kind == "solved":
        f = safediv_pos(1.0, r.tau)
        x, y, s = x * f, y * f, s * f
        res.status_val = ABIP_SOLVED_INACCURATE if inacc else ABIP_SOLVED
        res.status = "Solved/Inaccurate" if inacc else "Solved"
    elif kind == "indeterminate":
        x[:], y[:], s[:] = nan, nan, nan
        res.status_val, res.status = ABIP_INDETERMINATE, "Indeterminate"
    elif kind == "infeasible":
        y, s = y / r.bt_y_by_tau, s / r.bt_y_by_tau
        x[:] = nan
        res.status_val = -7 if inacc else ABIP_INFEASIBLE
        res.status = "Infeasible/Inaccurate" if inacc else "Infeasible"
    else:
        x = x * (-1 / r.ct_x_by_tau)
        y[:], s[:] = nan, nan
        res.status_val = -6 if inacc else ABIP_UNBOUNDED
        res.status = "Unbounded/Inaccurate" if inacc else "Unbounded"
    if s_.normalize:
        x = x / (w.E * w.sc_b)
        y = y / (w.D * w.sc_c)
        s = s * (w.E / (w.sc_c * s_.scale))
    res.x, res.y, res.s = x, y, s
    res.ipm_iter, res.admm_iter = i + 1, k + 1
    if kind == "solved":
        res.rel_gap, res.res_pri, res.res_dual = r.rel_gap, r.res_pri, r.res_dual
        res.pobj, res.dobj = r.ct_x_by_tau / r.tau, r.bt_y_by_tau / r.tau
    res.tot_cg_its, res.n_solves = w.p.tot_cg_its, w.p.n_solves
    return res


def solve(A: sp.csc_matrix, b, c, stgs: Settings | None = None, trace: bool = False) -> Result:
    """ABIP(main) -> init -> solve, src/abip.c:2056-2297, 2341-2422 (validate omitted: inputs are trusted)."""
    stgs = stgs or Settings()
    w = Work(A, b, c, stgs)
    r = Residuals()
    l = w.l
    tr = []
    k = 0
    status = 0
    for i in range(stgs.max_ipm_iters):
        spmin = min(w.sp, stgs.sparsity_ratio)
        if spmin > 0.5:
            inner_stopper = int(round(w.mu ** -0.35))
        elif spmin > 0.2:
            inner_stopper = int(round(w.mu ** -1))
        else:
            inner_stopper = stgs.max_admm_iters
        w.fre_old = 0
        w.u_avg[:] = 0
        w.v_avg[:] = 0
        w.u_sumcon[:] = 0
        w.v_sumcon[:] = 0
        if stgs.avg_criterion:
            w.u[:] = w.u_avgcon
            w.v[:] = w.v_avgcon
        for j in range(inner_stopper):
            w.u_prev[:] = w.u
            w.v_prev[:] = w.v
            its = project_lin_sys(w, k)
            if stgs.half_update:
                half_update_dual_vars(w)
                project_barrier_dual(w)
            else:
                project_barrier(w)
                update_dual_vars(w)
            restart_vars(w, j, k)
            compute_avg(w, j)
            k += 1
            q = iterate_Q_norm_resd(w, j)
            if trace:
                tr.append((i, j, k, w.mu, w.beta, its, q))
            if q < w.gamma * w.mu:
                if stgs.half_update:
                    w.v[w.v < 0] = 1e-6
                break
            if w.final_check:
                calc_residuals(w, r, i, k)
                status = has_converged(w, r, i, k)
                if status != 0 or k + 1 >= stgs.max_admm_iters or i + 1 >= stgs.max_ipm_iters:
                    res = get_solution(w, r, status, i, k)
                    res.trace = tr
                    return res
        if w.mu < stgs.eps:
            w.final_check = 1
        calc_residuals(w, r, i, k)
        status = has_converged(w, r, i, k)
        if status != 0 or k + 1 >= stgs.max_admm_iters:
            res = get_solution(w, r, status, i, k)
            res.trace = tr
            return res
        update_mu(w, r)
        reinitialize_vars(w, 0)
        if stgs.adaptive:
            reinitialize_vars(w, 1)
            w.beta = 1
            update_adapt_params(w, k)
            reinitialize_vars(w, 2)
    res = Result(status_val=status)
    res.trace = tr
    return res
