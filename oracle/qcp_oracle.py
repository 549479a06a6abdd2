"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference ABIP-QCP path (general QCP vtable).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.  Parity status: the
reference's only test (test/test_abip_install.m:32-43) has no assertions; this restatement is pinned against
outputs of the reference itself, compiled unmodified (stub MKL headers) into oracle/_ref/libabip_qcp_ref.so with
linsys_solver = 1 (tests/test_qcp_oracle.py, tests/golden/qcp_golden.json).  The reference's own pcg dispatch is
broken; `linsys="pcg"` below restates the unreachable-but-present n-space qcp_pcg
(linsys.c:725-851) with the engine's tolerance policy and is validated against the exact solve.

Citations are relative to /root/reference/src/abip-qcp.
"""
from __future__ import annotations

from dataclasses import dataclass, field
import math
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

MIN_SCALE, MAX_SCALE = 1e-3, 1e3  # source/qcp_config.c:2-3
SIGMA, GAMMA = 0.8, 1.6           # include/glbopts.h:31-32


@dataclass
class Settings:
    """source/util.c:203-255."""
    normalize: int = 1
    scale: float = 1.0
    rho_x: float = 1.0
    rho_y: float = 1e-6
    rho_tau: float = 1.0
    max_ipm_iters: int = 500
    max_admm_iters: int = 1000000
    eps_p: float = 1e-3
    eps_d: float = 1e-3
    eps_g: float = 1e-3
    eps_inf: float = 1e-3
    eps_unb: float = 1e-3
    err_dif: float = 0.0
    alpha: float = 1.8
    inner_check_period: int = 500
    outer_check_period: int = 1
    psi: float = 1.0
    origin_scaling: int = 1
    ruiz_scaling: int = 1
    pc_scaling: int = 0
    time_limit: float = float("inf")


def cone_list(K: dict):
    """Cone blocks in the fixed column order q -> rq -> f -> z -> l (include/abip.h:63-76): (kind, start, dim)."""
    out, pos = [], 0
    for d in K.get("q", []) or []:
        if d == 0:
            continue
        out.append(("q", pos, int(d)))
        pos += int(d)
    for d in K.get("rq", []) or []:
        if d < 3:
            continue  # abip.c:379-381 (count is not advanced for skipped cones)
        out.append(("rq", pos, int(d)))
        pos += int(d)
    for kind in ("f", "z", "l"):
        d = int(K.get(kind, 0) or 0)
        if d:
            out.append((kind, pos, d))
            pos += d
    return out


def scaling_data(A, Q, b, c, K, st: Settings):
    """source/qcp_config.c:91-491.  Returns scaled A, Q, b, c and D, E, sc_b, sc_c."""
    A = A.tocsc(copy=True).astype(np.float64)
    Q = Q.tocsc(copy=True).astype(np.float64) if Q is not None else None
    m, n = A.shape
    D_hat, E_hat = np.ones(m), np.ones(n)
    min_row, max_row = MIN_SCALE * math.sqrt(n), MAX_SCALE * math.sqrt(n)
    min_col, max_col = MIN_SCALE * math.sqrt(m), MAX_SCALE * math.sqrt(m)
    colA = np.repeat(np.arange(n), np.diff(A.indptr))
    colQ = np.repeat(np.arange(n), np.diff(Q.indptr)) if Q is not None else None

    def colred(M, vals, op, fill=0.0):
        out = np.full(n, fill)
        nz = np.diff(M.indptr) > 0
        out[nz] = op.reduceat(vals, M.indptr[:-1][nz])
        return out

    def cone_average(E):
        pos = 0
        for key in ("q", "rq"):
            for d in K.get(key, []) or []:
                if d > 0:
                    E[pos:pos + d] = np.sum(E[pos:pos + d]) / d  # linalg.c vec_mean
                pos += d
        return E

    def apply(E, D):
        D = D.copy()
        D[D < min_row] = 1.0
        D = np.minimum(D, max_row)
        E = E.copy()
        E[E < min_col] = 1.0
        E = np.minimum(E, max_col)
        A.data /= E[colA]
        if Q is not None:
            Q.data /= E[colQ]
            Q.data /= E[Q.indices]
        A.data /= D[A.indices]
        return E, D

    if st.ruiz_scaling:  # :158-262
        for _ in range(10):
            E1 = np.sqrt(colred(A, np.abs(A.data), np.maximum))
            E2 = np.sqrt(colred(Q, np.abs(Q.data), np.maximum)) if Q is not None else np.zeros(n)
            E = cone_average(np.maximum(E1, E2))
            D = np.zeros(m)
            np.maximum.at(D, A.indices, np.abs(A.data))
            E, D = apply(E, np.sqrt(D))
            E_hat *= E
            D_hat *= D
    if st.origin_scaling:  # :264-356
        E1 = np.sqrt(colred(A, A.data ** 2, np.add))
        E2 = np.sqrt(colred(Q, Q.data ** 2, np.add)) if Q is not None else np.zeros(n)
        E = cone_average(np.sqrt(np.maximum(E1, E2)))
        D = np.sqrt(np.sqrt(np.bincount(A.indices, weights=A.data ** 2, minlength=m)))
        E, D = apply(E, D)
        E_hat *= E
        D_hat *= D
    if st.pc_scaling:  # :358-452 (alpha_pc = 1)
        E1 = np.sqrt(colred(A, np.abs(A.data), np.add))
        E2 = np.sqrt(colred(Q, np.abs(Q.data), np.add)) if Q is not None else np.zeros(n)
        E = cone_average(np.maximum(E1, E2))
        D = np.sqrt(np.bincount(A.indices, weights=np.abs(A.data), minlength=m))
        E, D = apply(E, D)
        E_hat *= E
        D_hat *= D
    sc = math.sqrt(math.sqrt(float(c @ c) + float(b @ b)))  # :454-455 (norms of the *unscaled* b, c)
    b = b / D_hat
    c = c / E_hat
    if sc < MIN_SCALE:
        sc = 1.0
    elif sc > MAX_SCALE:
        sc = MAX_SCALE
    sc_b = sc_c = 1.0 / sc
    return A, Q, b * (sc_b * st.scale), c * (sc_c * st.scale), D_hat, E_hat, sc_b, sc_c


# ---- cone barrier proximal operators: source/cones.c ------------------------------------------------------
def positive_orthant_prox(t, lam):
    """cones.c:279-289."""
    t = np.asarray(t, dtype=np.float64)
    out = np.empty_like(t)
    pos = t >= 0
    out[pos] = (t[pos] + np.sqrt(t[pos] ** 2 + 4 * lam)) / 2
    tn = t[~pos]
    out[~pos] = 2 * lam / (-tn * (1 + np.sqrt(1 + 4 * lam / tn ** 2)))
    return out


def soc_prox(t, lam):
    """cones.c:130-161."""
    a, b = t[0], t[1:]
    bn = float(b @ b)
    out = np.empty_like(t)
    if abs(a) <= 1e-9:
        out[0] = math.sqrt(2 * lam + bn / 4)
        out[1:] = 0.5 * b
        return out
    r = 16 * a * a / (8 * lam - a * a + bn + math.sqrt((8 * lam - a * a + bn) ** 2 + 32 * a * a * lam))
    s1 = (r - math.sqrt(r * (r + 8))) / 2
    s2 = (r + math.sqrt(r * (r + 8))) / 2
    s = s2 if a > 0 else s1
    out[0] = (s + 2) * a / s
    out[1:] = b * ((s + 2) / (s + 4))
    return out


def rsoc_prox(t, lam, x_old0):
    """cones.c:169-248; the zeta_eta + zeta_nu == 0 branch reads the previous x[0] (:185)."""
    ze, zn, zx = t[0], t[1], t[2:]
    xn = float(zx @ zx)
    out = np.empty_like(t)
    if ze + zn == 0:
        out[1] = (-ze + math.sqrt(ze * ze + 4 * lam + xn)) / 2
        out[0] = x_old0 + ze
        out[2:] = 0.5 * zx
        return out
    dlt = 2 * ze * zn - xn
    big = 4 * (ze * ze + zn * zn + xn) / lam + 16
    if dlt < 0:
        g = -dlt / (2 * lam)
        w = (2 * (ze + zn) ** 2 / lam) / g / (1 + 4 / g + math.sqrt(1 + big / g / g))
    else:
        g = dlt / (2 * lam)
        w = g * (1 - 4 / g + math.sqrt(1 + big / g / g)) / 2
    if ze + zn > 0:
        s = (w + math.sqrt(w * (w + 4))) / 2
        out[0] = (ze * (s + 1) ** 2 + zn * (s + 1)) / (s * (s + 2))
        out[1] = (zn * (s + 1) ** 2 + ze * (s + 1)) / (s * (s + 2))
        out[2:] = zx * ((s + 1) / (s + 2))
    elif w > 10:
        s = 2 / (w + 2 + math.sqrt(w * (w + 4)))
        out[0] = (ze * s ** 2 + zn * s) / ((s - 1) * (s + 1))
        out[1] = (zn * s ** 2 + ze * s) / ((s - 1) * (s + 1))
        out[2:] = zx * (s / (s + 1))
    else:
        s = (w - math.sqrt(w * (w + 4))) / 2
        out[0] = (ze * (s + 1) ** 2 + zn * (s + 1)) / (s * (s + 2))
        out[1] = (zn * (s + 1) ** 2 + ze * (s + 1)) / (s * (s + 2))
        out[2:] = zx * ((s + 1) / (s + 2))
    return out


class Work:
    """init_work / scaling / linsys init / update_work / pre_calculate: source/abip.c:834-992."""

    def __init__(self, A, Q, b, c, K, st: Settings, linsys: str = "direct", pcg_rtol: float = 1e-10):
        self.st, self.K = st, K
        self.m, self.n = A.shape
        m, n = self.m, self.n
        self.l = m + n + 1
        self.cones = cone_list(K)
        self.nm_inf_b = float(np.max(np.abs(b))) if m else 0.0
        self.nm_inf_c = float(np.max(np.abs(c)))
        (self.A, self.Q, self.b, self.c, self.D, self.E, self.sc_b, self.sc_c) = scaling_data(A, Q, b, c, K, st)
        self.Acsr = self.A.tocsr()
        self.rho_dr = np.concatenate([np.full(m, st.rho_y), np.full(n, st.rho_x), [st.rho_tau]])  # qcp_config.c:26-36
        self.linsys = linsys
        self.pcg_rtol = pcg_rtol
        self.cg_iters = 0
        self.n_solves = 0
        if linsys == "direct":  # K of form_qcp_kkt (:699-748), solved exactly
            Qm = self.Q if self.Q is not None else sp.csc_matrix((n, n))
            Kmat = sp.bmat([[-st.rho_y * sp.identity(m), -self.A], [-self.A.T, Qm + st.rho_x * sp.identity(n)]]).tocsc()
            self.lu = spla.splu(Kmat)
        else:  # init_qcp_precon :754-780 (n-space Jacobi) and the engine's m-space Schur preconditioner
            Mv = np.asarray(self.A.multiply(self.A).sum(axis=0)).ravel() / st.rho_y
            if self.Q is not None:
                Mv = Mv + self.Q.diagonal()
            self.M = 1.0 / (Mv + st.rho_x)
            self.Hd = st.rho_x + (self.Q.diagonal() if self.Q is not None else np.zeros(n))
            self.q_is_diag = self.Q is None or (self.Q - sp.diags(self.Q.diagonal())).count_nonzero() == 0
            self.Ms = 1.0 / (st.rho_y + np.asarray(self.Acsr.multiply(self.Acsr).multiply(1.0 / self.Hd).sum(axis=1)).ravel())
            self.inner_its = 0
        self.sigma, self.gamma, self.mu, self.beta = SIGMA, GAMMA, 1.0, 1.0
        # update_work :912-992
        x = np.zeros(n)
        for kind, pos, d in self.cones:
            if kind == "q":
                x[pos] = 1
            elif kind == "rq":
                x[pos] = x[pos + 1] = 1
            elif kind == "l":
                x[pos:pos + d] = 1
        self.u = np.concatenate([np.zeros(m), x, [1.0]])
        self.v = self.u.copy()
        self.v_origin = np.zeros(self.l)
        self.u_t = np.zeros(self.l)
        self.rel_ut = np.zeros(self.l)
        # pre_calculate :886-910
        self.r = np.concatenate([-self.b, self.c])
        self.solve_linsys(self.r, None, -1)
        self.a = float(self.rho_dr[m + n] + (self.r * self.rho_dr[:m + n]) @ self.r)

    def mat_vec(self, x):
        """source/linsys.c:725-750: (R_x + Q + A' R_y^-1 A) x."""
        y = self.st.rho_x * x
        if self.Q is not None:
            y = y + self.Q @ x
        return y + self.A.T @ ((self.Acsr @ x) / self.st.rho_y)

    def pcg(self, b, x0, tol):
        """source/linsys.c:755-851 qcp_pcg (stop on |r|_inf < tol)."""
        n = self.n
        if x0 is None:
            r, x = b.copy(), np.zeros(n)
        else:
            r, x = b - self.mat_vec(x0), x0.copy()
        if np.max(np.abs(r)) < max(tol, 1e-12):
            return x, 0
        z = r * self.M
        ztr = float(z @ r)
        p = z.copy()
        its = 0
        for i in range(n * 4 + 50):
            Gp = self.mat_vec(p)
            alpha = ztr / float(p @ Gp)
            x += alpha * p
            r -= alpha * Gp
            its = i + 1
            if np.max(np.abs(r)) < tol:
                break
            z = r * self.M
            ztr_prev, ztr = ztr, float(z @ r)
            p = p * (ztr / ztr_prev) + z
        return x, its

    def hinv(self, v):
        """(Q + rho_x I)^-1 v: exact for diagonal Q, else Jacobi-PCG to 1e-13 relative (engine design, see
        DESIGN.md: the m-space Schur system is as well conditioned as the LP one, the n-space system of the
        reference's qcp_pcg has condition ~ 1/rho_y)."""
        if self.q_is_diag:
            return v / self.Hd
        x = v / self.Hd
        r = v - (self.st.rho_x * x + self.Q @ x)
        z = r / self.Hd
        p = z.copy()
        rz = float(r @ z)
        tol = 1e-13 * math.sqrt(float(v @ v))
        for _ in range(200):
            if math.sqrt(float(r @ r)) <= tol:
                break
            Hp = self.st.rho_x * p + self.Q @ p
            al = rz / float(p @ Hp)
            x += al * p
            r -= al * Hp
            z = r / self.Hd
            rz_new = float(r @ z)
            p = z + (rz_new / rz) * p
            rz = rz_new
            self.inner_its += 1
        return x

    def schur_solve(self, b, warm, it):
        """Engine path: eliminate x = H^-1(b_x + A'y), solve (rho_y I + A H^-1 A') y = b_y - A H^-1 b_x by PCG."""
        m, n, st = self.m, self.n, self.st
        hb = self.hinv(b[m:])
        rhs = b[:m] - self.Acsr @ hb
        op = lambda y: st.rho_y * y + self.Acsr @ self.hinv(self.A.T @ y)
        y = np.zeros(m) if warm is None else warm[:m].copy()
        r = rhs - op(y) if warm is not None else rhs.copy()
        # relative to |rhs|, but never more than 13 digits below the residual of the warm start (rhs = 0 otherwise asks for 0)
        tol = max(self.pcg_rtol * math.sqrt(float(rhs @ rhs)), 1e-13 * math.sqrt(float(r @ r)))
        its = 0
        if math.sqrt(float(r @ r)) > tol:
            z = r * self.Ms
            p = z.copy()
            rz = float(r @ z)
            for i in range(m * 2 + 50):
                Gp = op(p)
                if not float(p @ Gp) > 0.0:
                    break
                al = rz / float(p @ Gp)
                y += al * p
                r -= al * Gp
                its = i + 1
                if math.sqrt(float(r @ r)) < tol:
                    break
                z = r * self.Ms
                rz_new = float(r @ z)
                p = z + (rz_new / rz) * p
                rz = rz_new
        if it >= 0:
            self.cg_iters += its
        b[:m] = y
        b[m:] = hb + self.hinv(self.A.T @ y)
        return its

    def solve_linsys(self, b, warm, it):
        """solve_qcp_linsys, source/qcp_config.c:826-881: b (m+n) overwritten by the solution of
        [rho_y I, A; -A', Q + rho_x I] [y; x] = [b_y; b_x]."""
        m, n, st = self.m, self.n, self.st
        self.n_solves += 1
        if self.linsys == "direct":
            rhs = b.copy()
            rhs[:m] *= -1
            b[:] = self.lu.solve(rhs)
            return 0
        if self.linsys == "schur":
            return self.schur_solve(b, warm, it)
        b[m:] += self.A.T @ (b[:m] / st.rho_y)
        # engine tolerance policy (the reference passes error_ratio by mistake, qcp_config.c:852-855):
        # relative to the inf-norm of the reduced right-hand side
        tol = self.pcg_rtol * max(float(np.max(np.abs(b[m:]))), 1e-300)
        x, its = self.pcg(b[m:].copy(), None if warm is None else warm[m:m + n], tol)
        if it >= 0:
            self.cg_iters += its
        b[m:] = x
        b[:m] = (b[:m] - self.Acsr @ x) / st.rho_y
        return its


def projection(w: Work, it: int):
    """source/abip.c:186-254 (the branch taken whenever Q != NULL or linsys_solver != 3)."""
    m, n = w.m, w.n
    mu_ = (w.u[:m + n] + w.v[:m + n]) * w.rho_dr[:m + n]
    eta = w.rho_dr[m + n] * (w.u[m + n] + w.v[m + n])
    warm = w.u[:m + n] + w.u[m + n] * w.r  # :207-209 (used by pcg only)
    p = mu_.copy()
    w.solve_linsys(p, warm, it)
    tem = p * w.rho_dr[:m + n]
    bq = float(w.r @ mu_) - 2 * float(w.r @ tem) - eta
    cq = -float(p[m:] @ (w.Q @ p[m:])) if w.Q is not None else 0.0
    if it > 0:
        tau = (-bq + math.sqrt(max(0.0, bq * bq - 4 * w.a * cq))) / (2 * w.a)
    else:
        tau = 1.0
    w.u_t[m + n] = tau
    w.u_t[:m + n] = p - tau * w.r


def solve_barrier_subproblem(w: Work):
    """source/abip.c:326-413."""
    m, n, l, st = w.m, w.n, w.l, w.st
    lam = w.mu / w.beta
    w.rel_ut = st.alpha * w.u_t + (1 - st.alpha) * w.u - w.v
    tmp = w.rel_ut
    u_old = w.u.copy()
    w.u[:m] = tmp[:m]
    w.u[l - 1] = (tmp[l - 1] + math.sqrt(tmp[l - 1] ** 2 + 4 * lam / w.rho_dr[l - 1])) / 2
    for kind, pos, d in w.cones:
        lo, hi = m + pos, m + pos + d
        lc = lam / w.rho_dr[lo]
        if kind == "q":
            w.u[lo:hi] = positive_orthant_prox(tmp[lo:hi], lc) if d == 1 else soc_prox(tmp[lo:hi], lc)
        elif kind == "rq":
            w.u[lo:hi] = rsoc_prox(tmp[lo:hi], lc, u_old[lo])
        elif kind == "f":
            w.u[lo:hi] = tmp[lo:hi]
        elif kind == "z":
            w.u[lo:hi] = 0.0
        else:
            w.u[lo:hi] = positive_orthant_prox(tmp[lo:hi], lc)


def update_dual_vars(w: Work):
    """source/abip.c:314-324 and :1143-1144."""
    w.v = w.u - w.rel_ut
    w.v_origin = w.v * w.rho_dr


def inner_conv_check(w: Work):
    """source/qcp_config.c:518-557."""
    m, n = w.m, w.n
    y, x, tau = w.u[:m], w.u[m:m + n], w.u[m + n]
    Mu = np.concatenate([w.Acsr @ x, -(w.A.T @ y) + (w.Q @ x if w.Q is not None else 0.0)])
    Qu = np.empty(m + n + 1)
    Qu[:m] = Mu[:m] - tau * w.b
    Qu[m:m + n] = Mu[m:] + tau * w.c
    Qu[m + n] = -float(w.u[:m + n] @ Mu) / tau + float(y @ w.b) - float(x @ w.c)
    return float(np.linalg.norm(Qu - w.v_origin) / (1 + np.linalg.norm(Qu) + np.linalg.norm(w.v_origin)))


@dataclass
class Residuals:
    last_admm_iter: int = -1
    res_pri: float = 1e8
    res_dual: float = 1e8
    rel_gap: float = 1e8
    error_ratio: float = 1e8
    res_dif: float = float("nan")
    res_infeas: float = float("nan")
    res_unbdd: float = float("nan")
    pobj: float = float("nan")
    dobj: float = float("nan")
    tau: float = float("nan")
    kap: float = float("nan")
    Ax_b_norm: float = float("nan")
    Qx_ATy_c_s_norm: float = float("nan")


def calc_residuals(w: Work, r: Residuals, admm_iter: int):
    """source/qcp_config.c:562-691."""
    m, n, st = w.m, w.n, w.st
    if admm_iter and r.last_admm_iter == admm_iter:
        return
    r.last_admm_iter = admm_iter
    r.tau = abs(w.u[n + m])
    r.kap = abs(w.v_origin[n + m]) / ((st.scale * w.sc_c * w.sc_b) if st.normalize else 1.0)
    y, x, s = w.u[:m] / r.tau, w.u[m:m + n] / r.tau, w.v_origin[m:m + n] / r.tau
    Ax = w.Acsr @ x
    Ax_b = Ax - w.b
    r.Ax_b_norm = float(np.max(np.abs(Ax_b))) if m else 0.0
    this_pr = float(np.max(np.abs(Ax_b * w.D))) / (w.sc_b + max(float(np.max(np.abs(Ax * w.D))), w.sc_b * w.nm_inf_b))
    Qx = w.Q @ x if w.Q is not None else np.zeros(n)
    xQx_2 = float(x @ Qx) / (2 * w.sc_b * w.sc_c) if w.Q is not None else 0.0
    ATy = w.A.T @ y
    resd = Qx - ATy + w.c - s
    r.Qx_ATy_c_s_norm = float(np.max(np.abs(resd)))
    this_dr = float(np.max(np.abs(resd * w.E))) / (w.sc_c + max(w.sc_c * w.nm_inf_c, float(np.max(np.abs(Qx * w.E)))))
    cTx = float(w.c @ x) / (w.sc_b * w.sc_c)
    bTy = float(w.b @ y) / (w.sc_b * w.sc_c)
    this_gap = abs(2 * xQx_2 + cTx - bTy) / (1 + max(2 * xQx_2, max(abs(cTx), abs(bTy))))
    r.pobj, r.dobj = xQx_2 + cTx, -xQx_2 + bTy
    r.res_dif = max(abs(this_pr - r.res_pri), abs(this_dr - r.res_dual), abs(this_gap - r.rel_gap))
    r.res_pri, r.res_dual, r.rel_gap = this_pr, this_dr, this_gap
    r.error_ratio = max(r.res_pri / st.eps_p, r.res_dual / st.eps_d, r.rel_gap / st.eps_g)
    ctu = float(w.c @ w.u[m:m + n])
    r.res_unbdd = max(np.linalg.norm(Qx * w.E * r.tau), np.linalg.norm(Ax * w.D * r.tau)) / (-ctu) if ctu < 0 else math.inf
    btu = float(w.b @ w.u[:m])
    r.res_infeas = float(np.linalg.norm((ATy * w.E) * r.tau + (s * w.E) * r.tau)) / btu if btu > 0 else math.inf


def has_converged(w: Work, r: Residuals, ipm_iter: int, admm_iter: int) -> int:
    """source/abip.c:750-777."""
    st = w.st
    if r.res_pri < st.eps_p and r.res_dual < st.eps_d and r.rel_gap < st.eps_g:
        return 1
    if r.res_dif < st.err_dif * max(st.eps_p, st.eps_d, st.eps_g):
        return 2
    if r.res_unbdd < st.eps_unb and ipm_iter > 0 and admm_iter > 0:
        return -1
    if r.res_infeas < st.eps_inf and ipm_iter > 0 and admm_iter > 0:
        return -2
    return 0


def adjust_barrier(w: Work, r: Residuals) -> float:
    """source/abip.c:994-1071; returns the next tol_inner."""
    st = w.st
    sigma = 0.8
    ratio = w.mu / min(st.eps_p, st.eps_d, st.eps_g)
    gamma = 0.5
    for lo, hi, g in ((50, 100, 1.5), (10, 50, 1.3), (5, 10, 1.2), (1, 5, 1.1), (0.5, 1, 1.0), (0.1, 0.5, 0.9),
                      (0.05, 0.1, 0.9), (0.01, 0.05, 0.8), (0.005, 0.01, 0.8), (0.001, 0.005, 0.7),
                      (0.0005, 0.001, 0.7), (0.0001, 0.0005, 0.6), (0.00005, 0.0001, 0.6)):
        if lo < ratio <= hi:
            gamma = g
            break
    e = r.error_ratio
    if e > 22:
        gamma *= 4.4
    elif 18 < e <= 22:
        gamma *= 4.2
    elif 15 < e <= 18:
        gamma *= 4
    elif 12 < e <= 15:
        gamma *= 3.8
    elif 8 < e <= 12:
        gamma *= 3.6
    elif 6 < e <= 8:
        sigma, gamma = 0.81, gamma * 3.4
    elif 4 < e <= 6:
        sigma, gamma = 0.82, gamma * 3.4
    elif 3 < e <= 4:
        sigma, gamma = 0.83, gamma * 3.2
    elif 2 < e <= 3:
        sigma, gamma = 0.85, gamma * 2.8
    elif 1.5 < e <= 2:
        sigma, gamma = 0.85, gamma * 2.6
    elif e < 1.5:
        sigma, gamma = 0.85, gamma * 2.4
    sigma *= 0.2
    w.mu = sigma * w.mu
    return gamma * w.mu ** st.psi


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
    cg_iters: int = 0
    n_solves: int = 0
    trace: list = field(default_factory=list)


def get_solution(w: Work, r: Residuals, status_val: int, i: int, k: int) -> Result:
    """source/abip.c:559-586, 428-445, 530-557; un_scaling_qcp_sol qcp_config.c:496-513."""
    m, n = w.m, w.n
    x, y, s = w.u[m:m + n].copy(), w.u[:m].copy(), w.v[m:m + n].copy()
    res = Result()
    if status_val in (0, 1, 2):
        f = 1.0 / r.tau if r.tau >= 1e-18 else 1.0 / 1e-18
        x, y, s = x * f, y * f, s * f
        res.status_val = 2 if status_val in (0, 2) else 1
        res.status = "Solved/Inaccurate" if res.status_val == 2 else "Solved"
    elif status_val in (-2, -7):
        y, s = y / (r.dobj * r.tau), s / (r.dobj * r.tau)
        x[:] = np.nan
        res.status_val, res.status = -2, "Infeasible"
    else:
        x = x * (-1 / (r.pobj * r.tau))
        y[:], s[:] = np.nan, np.nan
        res.status_val, res.status = -1, "Unbounded"
    if w.st.normalize:
        x = x / (w.E * w.sc_b)
        y = y / (w.D * w.sc_c)
        s = s * (w.E / (w.sc_c * w.st.scale))
    res.x, res.y, res.s = x, y, s
    res.ipm_iter, res.admm_iter = i + 1, k
    if res.status_val in (1, 2):
        res.rel_gap, res.res_pri, res.res_dual, res.pobj, res.dobj = r.rel_gap, r.res_pri, r.res_dual, r.pobj, r.dobj
    res.cg_iters, res.n_solves = w.cg_iters, w.n_solves
    return res


def solve(A, Q, b, c, K, st: Settings | None = None, linsys: str = "direct", pcg_rtol: float = 1e-10,
          trace: bool = False) -> Result:
    """abip() -> ABIP(solve), source/abip.c:1076-1249, 1335-1371."""
    st = st or Settings()
    w = Work(A, Q, np.asarray(b, dtype=np.float64), np.asarray(c, dtype=np.float64), K, st, linsys, pcg_rtol)
    r = Residuals()
    sparsity = 1  # qcp_config.c:19-23: integer division makes this 1 for every sparse input
    tol_inner = 4 * w.mu ** st.psi
    k = 0
    tr = []
    status = 0
    for i in range(st.max_ipm_iters):
        for j in range(st.max_admm_iters):
            projection(w, k)
            solve_barrier_subproblem(w)
            update_dual_vars(w)
            k += 1
            err_inner = inner_conv_check(w)
            if trace:
                tr.append((i, j, k, w.mu, err_inner, tol_inner))
            if err_inner < tol_inner:
                break
            if (j + 1) % st.inner_check_period == 0 or r.error_ratio <= 8:
                calc_residuals(w, r, k)
                status = has_converged(w, r, i, k)
                if status != 0 or k + 1 >= st.max_admm_iters * st.max_ipm_iters or i + 1 >= st.max_ipm_iters:
                    res = get_solution(w, r, status, i, k)
                    res.trace = tr
                    return res
        if sparsity or (i + 1) % st.outer_check_period == 0:
            calc_residuals(w, r, k)
            status = has_converged(w, r, i, k)
            if status != 0 or k + 1 >= st.max_admm_iters * st.max_ipm_iters or i + 1 >= st.max_ipm_iters:
                res = get_solution(w, r, status, i, k)
                res.trace = tr
                return res
        tol_inner = adjust_barrier(w, r)
    res = Result(status_val=status)
    res.trace = tr
    return res
