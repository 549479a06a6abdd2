// qcp_engine.cu -- persistent cooperative kernels of the ABIP-QCP engine (general QCP vtable of the reference,
// src/abip-qcp/source/{abip.c, qcp_config.c, cones.c, linsys.c}).  sm_100a only; no CPU fallback.
//
// Linear system of the projection step (qcp_config.c:699-748, 826-881):
//     [ rho_y I      A        ] [y]   [b_y]
//     [  -A'     Q + rho_x I  ] [x] = [b_x]
// The reference's indirect path runs CG on the n-space normal equations (rho_x I + Q + A'A / rho_y) x = ...
// (qcp_pcg, linsys.c:755-851), whose condition number is ~ 1/rho_y = 1e6; that path does not converge and is in fact
// unreachable as shipped.  This engine eliminates x instead:
//     H = Q + rho_x I,   (rho_y I + A H^-1 A') y = b_y - A H^-1 b_x,   x = H^-1 (b_x + A'y),
// an m-space system as well conditioned as ABIP-LP's (rho_y I + AA'), solved by Jacobi-preconditioned CG to a
// relative residual (default 1e-8).  H^-1 is exact when Q is diagonal or absent and an inner Jacobi-PCG (1e-13)
// otherwise (1e-3 x the outer tolerance inside the operator of the outer CG).  With that accuracy the ADMM iteration counts
// equal those of the reference's direct (QDLDL) path.
#include "qcp_engine.h"

#include <cmath>
#include <cstring>
#include <vector>

#include "spmv_host.h"

struct QcpCtx {
    int m, n;
    Csr A, AT, Q;
    int has_q, q_diag;
    const double *b, *c, *D, *E, *Hd, *Ms, *Mn, *r;  // Mn: n-space Jacobi preconditioner of the reference (init_qcp_precon)
    double rho_y, rho_x, rho_tau, alpha, a_coef, rtol;
    int use_nspace;                                 // projection through the reference's n-space qcp_pcg (ABIP_GPU_QCP_NSPACE=1)
    const int *cone_start, *cone_dim, *cone_kind;  // SOC (kind 0) / RSOC (kind 1) blocks, x-index space
    int n_cones, cone_vars;                         // cone_vars = total variables in SOC/RSOC blocks
    int f_len, z_len, l_len;                        // then free, zero, orthant ranges (in this order)
    double *mu, *p, *warm, *hb;                     // rhs / solution [m+n], warm start [m], H^-1 b_x [n]
    double *cg_p, *cg_r, *cg_Gp, *tn1, *tn2;        // [m] x3, [n] x2
    double *ir, *ip, *iHp;                          // inner CG [n]
    double *partials, *sc;
};

struct QcpIterArgs {
    double *u, *v, *ut;
    long k;
    double mu_bar, beta;
};

// ---------------------------------------------------------------------------------------------------------
// x_out = (Q + rho_x I)^-1 v_in by Jacobi-preconditioned CG (general sparse symmetric Q).  Ends with all results
// visible grid-wide (last phase is followed by a grid barrier).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int dev_hinv_general(const QcpCtx& c, Reducer& R, cg::grid_group& grid, const double* v,
                                                double* x, double itol = 1e-13) {
    const int n = c.n;
    double s1[1] = {0.0};
    GRID_STRIDE(j, n) {
        const double vj = v[j];
        x[j] = vj / __ldg(c.Hd + j);
        s1[0] = fma(vj, vj, s1[0]);
    }
    R.block_store<1>(s1);
    grid_sync(grid);
    R.finish<1>(s1);
    const double tol = itol * sqrt(s1[0]);
    double s2[2] = {0.0, 0.0};
    spmv_rows(c.Q, x, R.ws, nullptr, [&](int row, double a) {
        const double rj = v[row] - fma(c.rho_x, x[row], a);
        const double zj = rj / __ldg(c.Hd + row);
        c.ir[row] = rj;
        c.ip[row] = zj;
        s2[0] = fma(rj, rj, s2[0]);
        s2[1] = fma(rj, zj, s2[1]);
    });
    R.block_store<2>(s2);
    grid_sync(grid);
    R.finish<2>(s2);
    double rr = s2[0], rz = s2[1];
    int its = 0;
    // Two grid barriers per iteration (as in the LP engine's PCG, lp_device.cuh: dev_solve_lin_sys): the epilogue of the Q
    // pass also reduces r.Hp, Hp.Hp, (M r).Hp, (M Hp).Hp, so |r - al Hp|^2 and (M r').r' -- the stopping test and beta --
    // are known right after alpha, and x, r, p are updated in ONE phase; (M r).r and |r|^2 of the current residual are
    // re-measured by the same epilogue, so the expanded forms never accumulate.
    while (sqrt(rr) > tol && its < 500) {
        double d[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        spmv_rows(c.Q, c.ip, R.ws, nullptr, [&](int row, double a) {
            const double pj = c.ip[row], rj = c.ir[row], hd = __ldg(c.Hd + row);
            const double hp = fma(c.rho_x, pj, a);
            c.iHp[row] = hp;
            const double zj = rj / hd, mg = hp / hd;
            d[0] = fma(pj, hp, d[0]);
            d[1] = fma(zj, hp, d[1]);
            d[2] = fma(mg, hp, d[2]);
            d[3] = fma(rj, hp, d[3]);
            d[4] = fma(hp, hp, d[4]);
            d[5] = fma(zj, rj, d[5]);
            d[6] = fma(rj, rj, d[6]);
        });
        R.block_store<7>(d);
        grid_sync(grid);
        R.finish<7>(d);
        rz = d[5];
        const double al = rz / d[0];
        const double rr_new = fma(al, fma(al, d[4], -2.0 * d[3]), d[6]);
        const double rz_new = fma(al, fma(al, d[2], -2.0 * d[1]), rz);
        const double be = rz_new / rz;
        rr = fmax(rr_new, 0.0);
        ++its;
        const bool stop = !(sqrt(rr) > tol);
        GRID_STRIDE(j, n) {
            const double pj = c.ip[j];
            x[j] = fma(al, pj, x[j]);
            const double rj = fma(-al, c.iHp[j], c.ir[j]);
            c.ir[j] = rj;
            if (!stop) c.ip[j] = fma(be, pj, rj / __ldg(c.Hd + j));
        }
        grid_sync(grid);
    }
    return its;
}

// ---------------------------------------------------------------------------------------------------------
// The reference's own indirect path, restated on the device (selected with linsys_solver = 3 in abip_qcp_gpu, and
// through abipgpu_qcp_solve_nspace): n-space operator mat_vec (source/linsys.c:725-750)
//     G x = rho_x x + Q x + A' ((A x) / rho_y),
// Jacobi preconditioner init_qcp_precon (qcp_config.c:754-780: Mn_j = 1 / (sum_i A_ij^2 / rho_y + Q_jj + rho_x), kernel
// k_qcp_mn) and qcp_pcg (linsys.c:755-851: stop on |r|_inf < tol, skip when the initial |r|_inf < max(tol, 1e-12)).
// Its condition number is ~ 1 / rho_y = 1e6, which is why the Schur path above is the default.
// bx [n]: right-hand side on entry, solution on exit; warm [n] or nullptr.  Work vectors: tn1 (x), ir (r), ip (p),
// iHp (G p), tn2 (rho_x p + Q p), cg_p (A p / rho_y, [m]).  Ends with a grid barrier.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_nspace_matvec(const QcpCtx& c, Reducer& R, cg::grid_group& grid, const double* x, double* y,
                                                  const double* dot_with, double* dot_out) {
    const int n = c.n;
    spmv_rows(c.A, x, R.ws, nullptr, [&](int row, double a) { c.cg_p[row] = a / c.rho_y; });
    if (c.has_q) {
        spmv_rows(c.Q, x, R.ws, nullptr, [&](int row, double a) { c.tn2[row] = fma(c.rho_x, x[row], a); });
    } else {
        GRID_STRIDE(j, n) c.tn2[j] = c.rho_x * x[j];
    }
    grid_sync(grid);
    double d[1] = {0.0};
    spmv_rows(c.AT, c.cg_p, R.ws, nullptr, [&](int row, double a) {
        const double g = c.tn2[row] + a;
        y[row] = g;
        if (dot_with) d[0] = fma(dot_with[row], g, d[0]);
    });
    if (dot_with) {
        R.block_store<1>(d);
        grid_sync(grid);
        R.finish<1>(d);
        *dot_out = d[0];
    } else {
        grid_sync(grid);
    }
}

__device__ __forceinline__ void dev_qcp_pcg_nspace(const QcpCtx& c, Reducer& R, cg::grid_group& grid, double* bx, const double* warm,
                                                   double tol, long max_iter, int& its_out, double& res_out) {
    const int n = c.n;
    double* x = c.tn1;
    double* r = c.ir;
    double* p = c.ip;
    double* Gp = c.iHp;
    // r = b - G x0, x = x0   (:771-786)
    if (warm) {
        dev_nspace_matvec(c, R, grid, warm, Gp, nullptr, nullptr);
        GRID_STRIDE(j, n) { r[j] = bx[j] - Gp[j]; x[j] = warm[j]; }
    } else {
        GRID_STRIDE(j, n) { r[j] = bx[j]; x[j] = 0.0; }
    }
    grid_sync(grid);
    double s[2] = {0.0, 0.0};  // [|r|_inf (max), z.r]
    GRID_STRIDE(j, n) {
        const double rj = r[j], zj = rj * __ldg(c.Mn + j);
        p[j] = zj;
        s[0] = fmax(s[0], fabs(rj));
        s[1] = fma(zj, rj, s[1]);
    }
    {
        double mx[1] = {s[0]}, sm[1] = {s[1]};
        R.block_store_max<1>(mx, 0);
        R.block_store<1>(sm, 1);
        grid_sync(grid);
        double o[2];
        R.finish<2, 1u>(o);
        s[0] = o[0];
        s[1] = o[1];
    }
    int its = 0;
    double rinf = s[0], ztr = s[1];
    if (!(rinf < fmax(tol, 1e-12))) {
        for (long i = 0; i < max_iter; ++i) {
            double pGp;
            dev_nspace_matvec(c, R, grid, p, Gp, p, &pGp);
            const double alpha = ztr / pGp;
            double t[2] = {0.0, 0.0};
            GRID_STRIDE(j, n) {
                x[j] = fma(alpha, p[j], x[j]);
                const double rj = fma(-alpha, Gp[j], r[j]);
                r[j] = rj;
                t[0] = fmax(t[0], fabs(rj));
                t[1] = fma(rj * __ldg(c.Mn + j), rj, t[1]);
            }
            double mx[1] = {t[0]}, sm[1] = {t[1]};
            R.block_store_max<1>(mx, 0);
            R.block_store<1>(sm, 1);
            grid_sync(grid);
            double o[2];
            R.finish<2, 1u>(o);
            its = (int)i + 1;
            rinf = o[0];
            if (rinf < tol) break;
            const double beta = o[1] / ztr;
            ztr = o[1];
            GRID_STRIDE(j, n) p[j] = fma(p[j], beta, r[j] * __ldg(c.Mn + j));
            grid_sync(grid);
        }
    }
    GRID_STRIDE(j, n) bx[j] = x[j];
    grid_sync(grid);
    its_out = its;
    res_out = rinf;
}

// solve_qcp_linsys with the n-space PCG (qcp_config.c:826-881): b_x += A'(b_y / rho_y); PCG on x; b_y = (b_y - A x) / rho_y.
// Tolerance: rtol * |b_x|_inf of the reduced right-hand side (the reference passes error_ratio there by mistake, :852-855).
__device__ __forceinline__ void dev_solve_nspace(const QcpCtx& c, Reducer& R, cg::grid_group& grid, double* vec, const double* warm_x,
                                                 double rtol, long max_iter, int& its, double& res) {
    const int m = c.m;
    double* by = vec;
    double* bx = vec + m;
    GRID_STRIDE(i, m) c.cg_r[i] = by[i] / c.rho_y;
    grid_sync(grid);
    double mx[1] = {0.0};
    spmv_rows(c.AT, c.cg_r, R.ws, nullptr, [&](int row, double a) {
        const double v = bx[row] + a;
        bx[row] = v;
        mx[0] = fmax(mx[0], fabs(v));
    });
    R.block_store_max<1>(mx);
    grid_sync(grid);
    R.finish<1, 1u>(mx);
    const double tol = rtol * fmax(mx[0], 1e-300);
    dev_qcp_pcg_nspace(c, R, grid, bx, warm_x, tol, max_iter, its, res);
    spmv_rows(c.A, bx, R.ws, nullptr, [&](int row, double a) { by[row] = (by[row] - a) / c.rho_y; });
    grid_sync(grid);
}

struct QcpSolveOut {
    int its, inner;
    double tol, res;
};

// ---------------------------------------------------------------------------------------------------------
// Schur-complement solve of the projection system; p = [b_y; b_x] on entry, [y; x] on exit.  warm: y warm start
// or nullptr.  Ends with a grid barrier.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_qcp_solve(const QcpCtx& c, Reducer& R, cg::grid_group& grid, double* p,
                                              const double* warm, double rtol, QcpSolveOut& out) {
    const int m = c.m, n = c.n;
    double* by = p;
    double* bx = p + m;
    const bool diag = c.q_diag != 0;
    int inner = 0;
    // 1. hb = H^-1 b_x
    if (diag) {
        GRID_STRIDE(j, n) c.hb[j] = bx[j] / __ldg(c.Hd + j);
        grid_sync(grid);
    } else {
        inner += dev_hinv_general(c, R, grid, bx, c.hb);
    }
    // 2. rhs = b_y - A hb (kept in cg_r), |rhs|^2; t = H^-1 A' warm
    double a1[1] = {0.0};
    spmv_rows(c.A, c.hb, R.ws, warm ? &c.AT : nullptr, [&](int row, double a) {
        const double v = by[row] - a;
        c.cg_r[row] = v;
        a1[0] = fma(v, v, a1[0]);
    });
    if (warm)
        spmv_rows(c.AT, warm, R.ws, nullptr, [&](int row, double a) { c.tn2[row] = diag ? a / __ldg(c.Hd + row) : a; });
    R.block_store<1>(a1);
    grid_sync(grid);
    R.finish<1>(a1);
    const double tol_rhs = rtol * sqrt(a1[0]);
    // H^-1 inside the operator of the outer CG: three digits below the outer target are enough (an operator error of
    // itol perturbs the attainable outer residual by ~ itol x iterations); the two solves that enter the solution
    // directly (hb above, x below) stay at 1e-13
    const double itol = fmin(1e-10, fmax(1e-13, 1e-3 * rtol));
    double* t1 = c.tn2;
    if (warm && !diag) {
        inner += dev_hinv_general(c, R, grid, c.tn2, c.tn1, itol);
        t1 = c.tn1;
    }
    // 3. r = rhs - (rho_y w + A t), y = w; z = Ms r; p = z
    double a2[2] = {0.0, 0.0};
    if (warm) {
        spmv_rows(c.A, t1, R.ws, &c.AT, [&](int row, double a) {
            const double wi = warm[row];
            const double ri = c.cg_r[row] - fma(c.rho_y, wi, a);
            const double zi = __ldg(c.Ms + row) * ri;
            c.cg_r[row] = ri;
            by[row] = wi;
            c.cg_p[row] = zi;
            a2[0] = fma(ri, ri, a2[0]);
            a2[1] = fma(zi, ri, a2[1]);
        });
    } else {
        GRID_STRIDE(i, m) {
            const double ri = c.cg_r[i];
            const double zi = __ldg(c.Ms + i) * ri;
            by[i] = 0.0;
            c.cg_p[i] = zi;
            a2[0] = fma(ri, ri, a2[0]);
            a2[1] = fma(zi, ri, a2[1]);
        }
    }
    R.block_store<2>(a2);
    grid_sync(grid);
    R.finish<2>(a2);
    double rn = sqrt(a2[0]);
    double ipzr = a2[1];
    int its = 0;
    const int max_its = 2 * m + 50;
    // The target is relative to |rhs|; when the reduced right-hand side vanishes (it is exactly 0 in the first iteration of
    // the SVM-QP program, abip_b200/svm.py) that target is 0 and CG would run on rounding noise until 0 / 0 turns the
    // iterates into NaN: never ask for more than 13 digits below the residual of the warm start.
    const double tol = fmax(tol_rhs, 1e-13 * rn);
    if (rn > tol) {
        for (int it = 0; it < max_its; ++it) {
            spmv_rows(c.AT, c.cg_p, R.ws, &c.A,
                      [&](int row, double a) { c.tn2[row] = diag ? a / __ldg(c.Hd + row) : a; });
            grid_sync(grid);
            t1 = c.tn2;
            if (!diag) {
                inner += dev_hinv_general(c, R, grid, c.tn2, c.tn1, itol);
                t1 = c.tn1;
            }
            double d1[1] = {0.0};
            spmv_rows(c.A, t1, R.ws, &c.AT, [&](int row, double a) {
                const double pi = c.cg_p[row];
                const double gp = fma(c.rho_y, pi, a);
                c.cg_Gp[row] = gp;
                d1[0] = fma(pi, gp, d1[0]);
            });
            R.block_store<1>(d1);
            grid_sync(grid);
            R.finish<1>(d1);
            if (!(d1[0] > 0.0)) break;  // p = 0 (or not a number): nothing left to do in this direction
            const double al = ipzr / d1[0];
            double d2[2] = {0.0, 0.0};
            GRID_STRIDE(i, m) {
                by[i] = fma(al, c.cg_p[i], by[i]);
                const double ri = fma(-al, c.cg_Gp[i], c.cg_r[i]);
                c.cg_r[i] = ri;
                const double zi = __ldg(c.Ms + i) * ri;
                d2[0] = fma(ri, ri, d2[0]);
                d2[1] = fma(zi, ri, d2[1]);
            }
            R.block_store<2>(d2);
            grid_sync(grid);
            R.finish<2>(d2);
            its = it + 1;
            rn = sqrt(d2[0]);
            if (rn < tol) break;
            const double be = d2[1] / ipzr;
            ipzr = d2[1];
            GRID_STRIDE(i, m) c.cg_p[i] = fma(be, c.cg_p[i], __ldg(c.Ms + i) * c.cg_r[i]);
            grid_sync(grid);
        }
    }
    // 5. x = hb + H^-1 A'y
    if (diag) {
        spmv_rows(c.AT, by, R.ws, &c.A, [&](int row, double a) { bx[row] = c.hb[row] + a / __ldg(c.Hd + row); });
        grid_sync(grid);
    } else {
        spmv_rows(c.AT, by, R.ws, nullptr, [&](int row, double a) { c.tn2[row] = a; });
        grid_sync(grid);
        inner += dev_hinv_general(c, R, grid, c.tn2, c.tn1);
        GRID_STRIDE(j, n) bx[j] = c.hb[j] + c.tn1[j];
        grid_sync(grid);
    }
    out.its = its;
    out.inner = inner;
    out.tol = tol;
    out.res = rn;
}

// ---- cone barrier proximal operators (src/abip-qcp/source/cones.c) -----------------------------------------
__device__ __forceinline__ double orthant_prox(double t, double lam) {  // cones.c:279-289
    if (t >= 0) return (t + sqrt(t * t + 4 * lam)) / 2;
    return 2 * lam / (-t * (1 + sqrt(1 + 4 * lam / (t * t))));
}

// rel_ut for one coordinate (abip.c:336-338) from the solve result p, r and the previous (u, v)
__device__ __forceinline__ double rel_at(const QcpCtx& c, const double* p, const double* u, const double* v, int i,
                                         double tau_t, double& ut_out) {
    const double ut = fma(-tau_t, __ldg(c.r + i), p[i]);
    ut_out = ut;
    return c.alpha * ut + (1 - c.alpha) * u[i] - v[i];
}

// One full inner ADMM iteration of ABIP-QCP (abip.c:1130-1152): projection, barrier subproblem, dual update and
// the sums of inner_conv_check (qcp_config.c:518-557) + calc_residuals (:562-691).
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_qcp_iter(QcpCtx c, QcpIterArgs a) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, c.partials);
    const int m = c.m, n = c.n, mn = m + n, lane = threadIdx.x & 31;
    double* u = a.u;
    double* v = a.v;
    double* p = c.p;
    const double utau = u[mn], vtau = v[mn];
    const double eta = c.rho_tau * (utau + vtau);

    // ---- projection (abip.c:186-254): rhs mu = rho o (u + v), warm start y = u_y + u_tau r_y
    spmv_prefetch(c.A, R.ws);
    double s1[1] = {0.0};
    GRID_STRIDE(i, mn) {
        const double ui = u[i];
        const double mui = (i < m ? c.rho_y : c.rho_x) * (ui + v[i]);
        c.mu[i] = mui;
        p[i] = mui;
        if (i < m) c.warm[i] = fma(utau, __ldg(c.r + i), ui);
        else if (c.use_nspace) c.hb[i - m] = fma(utau, __ldg(c.r + i), ui);  // x warm start u_x + tau r_x (abip.c:207-209)
        s1[0] = fma(__ldg(c.r + i), mui, s1[0]);
    }
    R.block_store<1>(s1);
    grid_sync(grid);
    R.finish<1>(s1);
    const double r_mu = s1[0];
    QcpSolveOut so;
    if (c.use_nspace) {
        so.inner = 0;
        so.tol = c.rtol;
        dev_solve_nspace(c, R, grid, p, c.hb, c.rtol, (long)c.n, so.its, so.res);
    } else {
        dev_qcp_solve(c, R, grid, p, c.warm, c.rtol, so);
    }
    // tau~ from a tau^2 + b tau + c = 0 (abip.c:228-246)
    double s2[2] = {0.0, 0.0};
    GRID_STRIDE(i, mn) s2[0] = fma(__ldg(c.r + i), (i < m ? c.rho_y : c.rho_x) * p[i], s2[0]);
    if (c.has_q)
        spmv_rows(c.Q, p + m, R.ws, nullptr, [&](int row, double q) { s2[1] = fma(p[m + row], q, s2[1]); });
    R.block_store<2>(s2);
    grid_sync(grid);
    R.finish<2>(s2);
    const double bq = r_mu - 2 * s2[0] - eta, cq = -s2[1];
    const double tau_t = a.k > 0 ? (-bq + sqrt(fmax(0.0, bq * bq - 4 * c.a_coef * cq))) / (2 * c.a_coef) : 1.0;

    // ---- solve_barrier_subproblem + update_dual_vars (abip.c:314-413)
    const double lam = a.mu_bar / a.beta;
    const double lam_x = lam / c.rho_x;
    GRID_STRIDE(i, m) {  // y block: u = rel_ut, v = 0
        double ut;
        const double rel = rel_at(c, p, u, v, i, tau_t, ut);
        a.ut[i] = ut;
        u[i] = rel;
        v[i] = rel - rel;
    }
    {   // free / zero / orthant ranges of x
        const int tail0 = c.cone_vars, tail = n - c.cone_vars;
        GRID_STRIDE(t, tail) {
            const int j = tail0 + t, i = m + j;
            double ut;
            const double rel = rel_at(c, p, u, v, i, tau_t, ut);
            double un;
            if (t < c.f_len) un = rel;
            else if (t < c.f_len + c.z_len) un = 0.0;
            else un = orthant_prox(rel, lam_x);
            a.ut[i] = ut;
            u[i] = un;
            v[i] = un - rel;
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // tau (abip.c:348-354)
        const double rel = c.alpha * tau_t + (1 - c.alpha) * utau - vtau;
        const double un = (rel + sqrt(rel * rel + 4 * lam / c.rho_tau)) / 2;
        a.ut[mn] = tau_t;
        u[mn] = un;
        v[mn] = un - rel;
    }
    {   // SOC / RSOC blocks: one warp per cone (cones.c:130-248)
        const int gwarp = blockIdx.x * kWarps + (threadIdx.x >> 5), nwarps = gridDim.x * kWarps;
        for (int ci = gwarp; ci < c.n_cones; ci += nwarps) {
            const int j0 = __ldg(c.cone_start + ci), d = __ldg(c.cone_dim + ci), kind = __ldg(c.cone_kind + ci);
            const int i0 = m + j0;
            const int nh = kind == 1 ? 2 : 1;  // head entries
            // pass 1: rel for the whole cone (kept in ut as scratch), |tail|^2
            double nsq = 0.0, h0 = 0.0, h1 = 0.0;
            for (int q = lane; q < d; q += 32) {
                double ut;
                const double rel = rel_at(c, p, u, v, i0 + q, tau_t, ut);
                a.ut[i0 + q] = ut;
                c.tn1[j0 + q] = rel;
                if (q >= nh) nsq = fma(rel, rel, nsq);
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) nsq += __shfl_xor_sync(0xffffffffu, nsq, off);
            __syncwarp();
            h0 = c.tn1[j0];
            if (nh == 2) h1 = c.tn1[j0 + 1];
            double x0, x1 = 0.0, sc_tail;
            if (kind == 0 && d == 1) {  // abip.c:363-366
                x0 = orthant_prox(h0, lam_x);
                sc_tail = 0.0;
            } else if (kind == 0) {  // soc_barrier_subproblem
                const double aa = h0;
                if (fabs(aa) <= 1e-9) {
                    x0 = sqrt(2 * lam_x + nsq / 4);
                    sc_tail = 0.5;
                } else {
                    const double w8 = 8 * lam_x - aa * aa + nsq;
                    const double rr = 16 * aa * aa / (w8 + sqrt(w8 * w8 + 32 * aa * aa * lam_x));
                    const double sq = sqrt(rr * (rr + 8));
                    const double s = aa > 0 ? (rr + sq) / 2 : (rr - sq) / 2;
                    x0 = (s + 2) * aa / s;
                    sc_tail = (s + 2) / (s + 4);
                }
            } else {  // rsoc_barrier_subproblem
                const double ze = h0, zn = h1;
                if (ze + zn == 0) {
                    x1 = (-ze + sqrt(ze * ze + 4 * lam_x + nsq)) / 2;
                    x0 = u[i0] + ze;  // reads the previous x[0] (cones.c:185)
                    sc_tail = 0.5;
                } else {
                    const double dlt = 2 * ze * zn - nsq;
                    const double big = 4 * (ze * ze + zn * zn + nsq) / lam_x + 16;
                    double w;
                    if (dlt < 0) {
                        const double g = -dlt / (2 * lam_x);
                        w = (2 * (ze + zn) * (ze + zn) / lam_x) / g / (1 + 4 / g + sqrt(1 + big / g / g));
                    } else {
                        const double g = dlt / (2 * lam_x);
                        w = g * (1 - 4 / g + sqrt(1 + big / g / g)) / 2;
                    }
                    if (ze + zn > 0) {
                        const double s = (w + sqrt(w * (w + 4))) / 2;
                        x0 = (ze * (s + 1) * (s + 1) + zn * (s + 1)) / (s * (s + 2));
                        x1 = (zn * (s + 1) * (s + 1) + ze * (s + 1)) / (s * (s + 2));
                        sc_tail = (s + 1) / (s + 2);
                    } else if (w > 10) {
                        const double s = 2 / (w + 2 + sqrt(w * (w + 4)));
                        x0 = (ze * s * s + zn * s) / ((s - 1) * (s + 1));
                        x1 = (zn * s * s + ze * s) / ((s - 1) * (s + 1));
                        sc_tail = s / (s + 1);
                    } else {
                        const double s = (w - sqrt(w * (w + 4))) / 2;
                        x0 = (ze * (s + 1) * (s + 1) + zn * (s + 1)) / (s * (s + 2));
                        x1 = (zn * (s + 1) * (s + 1) + ze * (s + 1)) / (s * (s + 2));
                        sc_tail = (s + 1) / (s + 2);
                    }
                }
            }
            __syncwarp();
            for (int q = lane; q < d; q += 32) {
                const double rel = c.tn1[j0 + q];
                const double un = q == 0 ? x0 : ((q == 1 && nh == 2) ? x1 : rel * sc_tail);
                u[i0 + q] = un;
                v[i0 + q] = un - rel;
            }
        }
    }
    grid_sync(grid);

    // ---- inner_conv_check + calc_residuals sums on the new (u, v)
    const double tau = u[mn];
    const double itau = 1.0 / fabs(tau);
    double sA[5] = {0, 0, 0, 0, 0};  // S_DIFF_y, S_QU_y, UMU_y, YB, AXD2
    double mA[3] = {0, 0, 0};        // AXB_INF, AXB_D_INF, AX_D_INF
    spmv_rows(c.A, u + m, R.ws, c.has_q ? &c.Q : &c.AT, [&](int row, double ax) {
        const double bi = __ldg(c.b + row), di = __ldg(c.D + row), yi = u[row];
        const double qu = fma(-tau, bi, ax);
        const double vo = c.rho_y * v[row];
        sA[0] = fma(qu - vo, qu - vo, sA[0]);
        sA[1] = fma(qu, qu, sA[1]);
        sA[2] = fma(yi, ax, sA[2]);
        sA[3] = fma(yi, bi, sA[3]);
        sA[4] = fma(di * ax, di * ax, sA[4]);
        const double axb = ax * itau - bi;
        mA[0] = fmax(mA[0], fabs(axb));
        mA[1] = fmax(mA[1], fabs(axb * di));
        mA[2] = fmax(mA[2], fabs(ax * itau * di));
    });
    if (c.has_q) spmv_rows(c.Q, u + m, R.ws, &c.AT, [&](int row, double q) { c.tn1[row] = q; });
    // (tn1 rows are produced and consumed by the same warp partition only if the plans coincide -- they do not,
    //  so a barrier is required between the Q pass and the A' pass)
    if (c.has_q) grid_sync(grid);
    double sX[7] = {0, 0, 0, 0, 0, 0, 0};  // S_DIFF_x, S_QU_x, S_VO_x, UMU_x, XC, XQX, QXE2 | ATYS_E2 below
    double sX2[1] = {0};
    double mX[3] = {0, 0, 0};  // RESD_INF, RESD_E_INF, QX_E_INF
    spmv_rows(c.AT, u, R.ws, nullptr, [&](int row, double aty) {
        const double cj = __ldg(c.c + row), ej = __ldg(c.E + row), xj = u[m + row];
        const double qx = c.has_q ? c.tn1[row] : 0.0;
        const double mux = qx - aty;
        const double qu = fma(tau, cj, mux);
        const double vo = c.rho_x * v[m + row];
        sX[0] = fma(qu - vo, qu - vo, sX[0]);
        sX[1] = fma(qu, qu, sX[1]);
        sX[2] = fma(vo, vo, sX[2]);
        sX[3] = fma(xj, mux, sX[3]);
        sX[4] = fma(xj, cj, sX[4]);
        sX[5] = fma(xj, qx, sX[5]);
        sX[6] = fma(ej * qx, ej * qx, sX[6]);
        const double ae = ej * (aty + vo);
        sX2[0] = fma(ae, ae, sX2[0]);
        const double resd = (mux - vo) * itau + cj;
        mX[0] = fmax(mX[0], fabs(resd));
        mX[1] = fmax(mX[1], fabs(resd * ej));
        mX[2] = fmax(mX[2], fabs(qx * itau * ej));
    });
    R.ws.drain();
    // slots: 0..4 sA, 5..11 sX, 12 sX2, 13..15 mA, 16..18 mX
    R.block_store<5>(sA, 0);
    R.block_store<7>(sX, 5);
    R.block_store<1>(sX2, 12);
    R.block_store_max<3>(mA, 13);
    R.block_store_max<3>(mX, 16);
    grid_sync(grid);
    double t[19];
    R.finish<19, 0x7E000u>(t);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double* sc = c.sc;
        sc[ABIPGPU_QSC_CG_ITS] = (double)so.its;
        sc[ABIPGPU_QSC_INNER_ITS] = (double)so.inner;
        sc[ABIPGPU_QSC_TAU_T] = tau_t;
        sc[ABIPGPU_QSC_S_DIFF] = t[0] + t[5];
        sc[ABIPGPU_QSC_S_QU] = t[1] + t[6];
        sc[ABIPGPU_QSC_S_VO] = t[7];  // v_y == 0: only the x block contributes
        sc[ABIPGPU_QSC_UMU] = t[2] + t[8];
        sc[ABIPGPU_QSC_YB] = t[3];
        sc[ABIPGPU_QSC_XC] = t[9];
        sc[ABIPGPU_QSC_XQX] = t[10];
        sc[ABIPGPU_QSC_AXD2] = t[4];
        sc[ABIPGPU_QSC_QXE2] = t[11];
        sc[ABIPGPU_QSC_ATYS_E2] = t[12];
        sc[ABIPGPU_QSC_AXB_INF] = t[13];
        sc[ABIPGPU_QSC_AXB_D_INF] = t[14];
        sc[ABIPGPU_QSC_AX_D_INF] = t[15];
        sc[ABIPGPU_QSC_RESD_INF] = t[16];
        sc[ABIPGPU_QSC_RESD_E_INF] = t[17];
        sc[ABIPGPU_QSC_QX_E_INF] = t[18];
        sc[ABIPGPU_QSC_TAU] = tau;
        sc[ABIPGPU_QSC_VO_TAU] = c.rho_tau * v[mn];
        sc[ABIPGPU_QSC_CG_RES] = so.res;
    }
}

__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM)
    k_qcp_solve_nspace(QcpCtx c, double* vec, const double* warm_x, double rtol, long max_iter) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, c.partials);
    int its;
    double res;
    dev_solve_nspace(c, R, grid, vec, warm_x, rtol, max_iter, its, res);
    R.ws.drain();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_QSC_CG_ITS] = (double)its;
        c.sc[ABIPGPU_QSC_INNER_ITS] = 0.0;
        c.sc[ABIPGPU_QSC_CG_RES] = res;
    }
}

// init_qcp_precon (qcp_config.c:754-780): Mn_j = 1 / (sum_i A_ij^2 / rho_y + Q_jj + rho_x); AT = CSR(A') (rows = columns of A)
__global__ void k_qcp_mn(Csr AT, const double* Hd, double rho_y, double* Mn) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= AT.nrows) return;
    double s = 0.0;
    for (int k = AT.ptr[j]; k < AT.ptr[j + 1]; ++k) s = fma(AT.val[k], AT.val[k], s);
    Mn[j] = 1.0 / (s / rho_y + Hd[j]);
}

// pre_calculate (abip.c:886-910): r = K^-1 [-b; c] (tight tolerance), a = rho_tau + r'(rho o r)
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_qcp_precalc(QcpCtx c, double* rvec) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, c.partials);
    const int m = c.m, mn = c.m + c.n;
    GRID_STRIDE(i, mn) rvec[i] = i < m ? -__ldg(c.b + i) : __ldg(c.c + i - m);
    grid_sync(grid);
    QcpSolveOut so;
    dev_qcp_solve(c, R, grid, rvec, nullptr, 1e-12, so);
    R.ws.drain();
    double s[1] = {0.0};
    GRID_STRIDE(i, mn) s[0] = fma((i < m ? c.rho_y : c.rho_x) * rvec[i], rvec[i], s[0]);
    R.block_store<1>(s);
    grid_sync(grid);
    R.finish<1>(s);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_QSC_A_COEF] = c.rho_tau + s[0];
        c.sc[ABIPGPU_QSC_CG_ITS] = (double)so.its;
        c.sc[ABIPGPU_QSC_INNER_ITS] = (double)so.inner;
    }
}

// generic solve on a device vector (tests): vec = K^-1 vec
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM)
    k_qcp_solve_vec(QcpCtx c, double* vec, const double* warm, double rtol) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, c.partials);
    QcpSolveOut so;
    dev_qcp_solve(c, R, grid, vec, warm, rtol, so);
    R.ws.drain();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_QSC_CG_ITS] = (double)so.its;
        c.sc[ABIPGPU_QSC_INNER_ITS] = (double)so.inner;
        c.sc[ABIPGPU_QSC_CG_RES] = so.res;
    }
}

// Hd = rho_x + diag(Q); Ms = 1 / (rho_y + sum_j A_ij^2 / Hd_j)
__global__ void k_qcp_hd(Csr Q, int has_q, double rho_x, double* Hd, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double d = 0.0;
    if (has_q)
        for (int k = Q.ptr[j]; k < Q.ptr[j + 1]; ++k)
            if (Q.idx[k] == j) d += Q.val[k];
    Hd[j] = rho_x + d;
}
__global__ void k_qcp_ms(Csr A, const double* Hd, double rho_y, double* Ms) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= A.nrows) return;
    double s = 0.0;
    for (int k = A.ptr[i]; k < A.ptr[i + 1]; ++k) s += A.val[k] * A.val[k] / Hd[A.idx[k]];
    Ms[i] = 1.0 / (rho_y + s);
}

// =========================================================================================================
// Host side
// =========================================================================================================
struct ABIPGPU_QCP {
    int device = 0, m = 0, n = 0, l = 0, num_sms = 0, grid = 0;
    cudaStream_t stream = nullptr;
    DevCsr A, AT, Q;
    int has_q = 0, q_diag = 1;
    int *d_cone_start = nullptr, *d_cone_dim = nullptr, *d_cone_kind = nullptr;
    double* slab = nullptr;
    double *u = nullptr, *v = nullptr, *ut = nullptr, *r = nullptr;
    double* hsc = nullptr;
    QcpCtx ctx;
    long n_iter = 0, n_cg = 0, n_inner = 0;
    double kernel_ms = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

template <class... Args>
static int qlaunch(abipgpu_qcp* e, const void* kernel, Args... args) {
    void* argv[] = {(void*)&args...};
    CK(cudaLaunchCooperativeKernel(kernel, dim3(e->grid), dim3(kBlock), argv, kSmemBytes, e->stream));
    return 0;
}

static int qread_sc(abipgpu_qcp* e, double* sc) {
    CK(cudaMemcpyAsync(e->hsc, e->ctx.sc, sizeof(double) * ABIPGPU_QSC_COUNT, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    if (sc) memcpy(sc, e->hsc, sizeof(double) * ABIPGPU_QSC_COUNT);
    return 0;
}

static int qcp_create_impl(abipgpu_qcp* e, int m, int n, const int* Ap, const int* Ai, const double* Ax, const int* Qp,
                           const int* Qi, const double* Qx, const double* b, const double* c, const double* D,
                           const double* E, const int* q, int qsize, const int* rq, int rqsize, int f, int z, int l,
                           double rho_x, double rho_y, double rho_tau, double alpha, double rtol, int device) {
    e->device = device;
    e->m = m;
    e->n = n;
    e->l = m + n + 1;
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (!prop.cooperativeLaunch) return -1;
    e->num_sms = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&e->ev0));
    CK(cudaEventCreate(&e->ev1));
    {
        int occ[4];
        const void* ks[4] = {(const void*)k_qcp_iter, (const void*)k_qcp_precalc, (const void*)k_qcp_solve_vec,
                             (const void*)k_qcp_solve_nspace};
        int g = 1 << 30;
        for (int i = 0; i < 4; ++i) {
            CK(cudaFuncSetAttribute(ks[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[i], ks[i], kBlock, kSmemBytes));
            if (occ[i] < 1) {
                fprintf(stderr, "[abip_gpu] QCP kernel cannot be made resident\n");
                return -1;
            }
            g = std::min(g, e->num_sms * std::min(occ[i], ABIP_MIN_BLOCKS_PER_SM));
        }
        e->grid = g;
    }
    const int W = e->grid * kWarps;
    // CSR(A') = CSC(A); CSR(A) by transposition; Q symmetric: CSR(Q) = CSC(Q)
    std::vector<int> at_ptr(Ap, Ap + n + 1), at_idx(Ai, Ai + Ap[n]);
    std::vector<double> at_val(Ax, Ax + Ap[n]);
    std::vector<int> a_ptr, a_idx;
    std::vector<double> a_val;
    if (csc_to_csr<int>(m, n, Ap, Ai, Ax, &a_ptr, &a_idx, &a_val)) return -1;
    if (upload_csr(&e->A, a_ptr, a_idx, a_val, m, W, "ABIP_GPU_LANES_A", e->stream) ||
        upload_csr(&e->AT, at_ptr, at_idx, at_val, n, W, "ABIP_GPU_LANES_AT", e->stream))
        return -1;
    e->has_q = (Qp && Qp[n] > 0) ? 1 : 0;
    e->q_diag = 1;
    if (e->has_q) {
        std::vector<int> q_ptr(Qp, Qp + n + 1), q_idx(Qi, Qi + Qp[n]);
        std::vector<double> q_val(Qx, Qx + Qp[n]);
        for (int j = 0; j < n && e->q_diag; ++j)
            for (int k = Qp[j]; k < Qp[j + 1]; ++k)
                if (Qi[k] != j && Qx[k] != 0.0) { e->q_diag = 0; break; }
        if (upload_csr(&e->Q, q_ptr, q_idx, q_val, n, W, "ABIP_GPU_LANES_Q", e->stream)) return -1;
    } else {
        std::vector<int> q_ptr(n + 1, 0), q_idx;
        std::vector<double> q_val;
        if (upload_csr(&e->Q, q_ptr, q_idx, q_val, n, W, "ABIP_GPU_LANES_Q", e->stream)) return -1;
    }
    // cones (column order q -> rq -> f -> z -> l, include/abip.h:63-76)
    std::vector<int> cs, cd, ck;
    int pos = 0;
    for (int i = 0; i < qsize; ++i) {
        if (q[i] == 0) continue;
        cs.push_back(pos); cd.push_back(q[i]); ck.push_back(0);
        pos += q[i];
    }
    for (int i = 0; i < rqsize; ++i) {
        if (rq[i] < 3) continue;  // abip.c:379-381
        cs.push_back(pos); cd.push_back(rq[i]); ck.push_back(1);
        pos += rq[i];
    }
    if (pos + f + z + l != n) {
        fprintf(stderr, "[abip_gpu] cone dimensions %d do not match n = %d\n", pos + f + z + l, n);
        return -1;
    }
    if (upload_padded(&e->d_cone_start, cs, e->stream) || upload_padded(&e->d_cone_dim, cd, e->stream) ||
        upload_padded(&e->d_cone_kind, ck, e->stream))
        return -1;

    const size_t L = ((size_t)e->l + 31) & ~(size_t)31, Mm = ((size_t)m + 31) & ~(size_t)31,
                 Nn = ((size_t)n + 31) & ~(size_t)31;
    const size_t total = 7 * L + 8 * Mm + 10 * Nn + (size_t)2 * kMaxRed * e->grid + 64 + ABIPGPU_QSC_COUNT;
    CK(cudaMalloc((void**)&e->slab, total * sizeof(double)));
    CK(cudaMemsetAsync(e->slab, 0, total * sizeof(double), e->stream));
    double* qq = e->slab;
    auto take = [&](size_t k) { double* rr = qq; qq += k; return rr; };
    e->u = take(L); e->v = take(L); e->ut = take(L); e->r = take(L);
    double* mu = take(L); double* p = take(L); take(L);
    double* db = take(Mm); double* dD = take(Mm); double* Ms = take(Mm); double* warm = take(Mm);
    double* cg_p = take(Mm); double* cg_r = take(Mm); double* cg_Gp = take(Mm); take(Mm);
    double* dc = take(Nn); double* dE = take(Nn); double* Hd = take(Nn); double* hb = take(Nn);
    double* tn1 = take(Nn); double* tn2 = take(Nn); double* ir = take(Nn); double* ip = take(Nn);
    double* iHp = take(Nn); double* Mn = take(Nn);
    double* partials = take((size_t)2 * kMaxRed * e->grid + 64);
    double* dsc = take(ABIPGPU_QSC_COUNT);
    CK(cudaMallocHost((void**)&e->hsc, sizeof(double) * ABIPGPU_QSC_COUNT));
    CK(cudaMemcpyAsync(db, b, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dc, c, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dD, D, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(dE, E, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));

    QcpCtx& x = e->ctx;
    x.m = m; x.n = n;
    x.A = e->A.view(); x.AT = e->AT.view(); x.Q = e->Q.view();
    x.has_q = e->has_q; x.q_diag = e->q_diag;
    x.b = db; x.c = dc; x.D = dD; x.E = dE; x.Hd = Hd; x.Ms = Ms; x.Mn = Mn; x.r = e->r;
    x.rho_y = rho_y; x.rho_x = rho_x; x.rho_tau = rho_tau; x.alpha = alpha; x.a_coef = 0.0; x.rtol = rtol;
    x.use_nspace = env_int("ABIP_GPU_QCP_NSPACE", 0) != 0 ? 1 : 0;
    x.cone_start = e->d_cone_start; x.cone_dim = e->d_cone_dim; x.cone_kind = e->d_cone_kind;
    x.n_cones = (int)cs.size(); x.cone_vars = pos; x.f_len = f; x.z_len = z; x.l_len = l;
    x.mu = mu; x.p = p; x.warm = warm; x.hb = hb;
    x.cg_p = cg_p; x.cg_r = cg_r; x.cg_Gp = cg_Gp; x.tn1 = tn1; x.tn2 = tn2; x.ir = ir; x.ip = ip; x.iHp = iHp;
    x.partials = partials; x.sc = dsc;

    k_qcp_hd<<<(n + 255) / 256, 256, 0, e->stream>>>(x.Q, e->has_q, rho_x, Hd, n);
    k_qcp_ms<<<(m + 255) / 256, 256, 0, e->stream>>>(x.A, Hd, rho_y, Ms);
    k_qcp_mn<<<(n + 255) / 256, 256, 0, e->stream>>>(x.AT, Hd, rho_y, Mn);
    CK(cudaGetLastError());
    // update_work (abip.c:912-992): initial point
    std::vector<double> u0(e->l, 0.0);
    for (size_t ci = 0; ci < cs.size(); ++ci) {
        u0[m + cs[ci]] = 1.0;
        if (ck[ci] == 1) u0[m + cs[ci] + 1] = 1.0;
    }
    for (int j = 0; j < l; ++j) u0[m + pos + f + z + j] = 1.0;
    u0[m + n] = 1.0;
    CK(cudaMemcpyAsync(e->u, u0.data(), sizeof(double) * e->l, cudaMemcpyHostToDevice, e->stream));
    CK(cudaMemcpyAsync(e->v, u0.data(), sizeof(double) * e->l, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    // pre_calculate
    if (qlaunch(e, (const void*)k_qcp_precalc, e->ctx, e->r)) return -1;
    if (qread_sc(e, nullptr)) return -1;
    e->ctx.a_coef = e->hsc[ABIPGPU_QSC_A_COEF];
    return 0;
}

extern "C" {

abipgpu_qcp* abipgpu_qcp_create(int m, int n, const int* Ap, const int* Ai, const double* Ax, const int* Qp, const int* Qi,
                                const double* Qx, const double* b, const double* c, const double* D, const double* E,
                                const int* q, int qsize, const int* rq, int rqsize, int f, int z, int l, double rho_x,
                                double rho_y, double rho_tau, double alpha, double rtol, int device) {
    abipgpu_qcp* e = new abipgpu_qcp();
    if (qcp_create_impl(e, m, n, Ap, Ai, Ax, Qp, Qi, Qx, b, c, D, E, q, qsize, rq, rqsize, f, z, l, rho_x, rho_y, rho_tau,
                        alpha, rtol, device) != 0) {
        abipgpu_qcp_destroy(e);
        return nullptr;
    }
    return e;
}

void abipgpu_qcp_destroy(abipgpu_qcp* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->A.release(); e->AT.release(); e->Q.release();
    cudaFree(e->d_cone_start); cudaFree(e->d_cone_dim); cudaFree(e->d_cone_kind);
    cudaFree(e->slab);
    if (e->hsc) cudaFreeHost(e->hsc);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int abipgpu_qcp_iter(abipgpu_qcp* e, long k, double mu, double beta, double* sc) {
    CK(cudaSetDevice(e->device));
    QcpIterArgs a{e->u, e->v, e->ut, k, mu, beta};
    CK(cudaEventRecord(e->ev0, e->stream));
    if (qlaunch(e, (const void*)k_qcp_iter, e->ctx, a)) return -1;
    CK(cudaEventRecord(e->ev1, e->stream));
    if (qread_sc(e, sc)) return -1;
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
    e->kernel_ms += ms;
    e->n_iter++;
    e->n_cg += (long)e->hsc[ABIPGPU_QSC_CG_ITS];
    e->n_inner += (long)e->hsc[ABIPGPU_QSC_INNER_ITS];
    return 0;
}

int abipgpu_qcp_solve_vec(abipgpu_qcp* e, double* host_vec, const double* host_warm, double rtol, double* sc) {
    CK(cudaSetDevice(e->device));
    const int mn = e->m + e->n;
    CK(cudaMemcpyAsync(e->ctx.mu, host_vec, sizeof(double) * mn, cudaMemcpyHostToDevice, e->stream));
    if (host_warm) CK(cudaMemcpyAsync(e->ctx.warm, host_warm, sizeof(double) * e->m, cudaMemcpyHostToDevice, e->stream));
    if (qlaunch(e, (const void*)k_qcp_solve_vec, e->ctx, e->ctx.mu, (const double*)(host_warm ? e->ctx.warm : nullptr), rtol))
        return -1;
    CK(cudaMemcpyAsync(host_vec, e->ctx.mu, sizeof(double) * mn, cudaMemcpyDeviceToHost, e->stream));
    return qread_sc(e, sc);
}

// the reference's n-space path (mat_vec + init_qcp_precon + qcp_pcg + solve_qcp_linsys): vec [m+n] in/out, warm_x [n] or NULL
int abipgpu_qcp_solve_nspace(abipgpu_qcp* e, double* host_vec, const double* host_warm_x, double rtol, long max_iter, double* sc) {
    CK(cudaSetDevice(e->device));
    const int mn = e->m + e->n;
    CK(cudaMemcpyAsync(e->ctx.mu, host_vec, sizeof(double) * mn, cudaMemcpyHostToDevice, e->stream));
    if (host_warm_x) CK(cudaMemcpyAsync(e->ctx.hb, host_warm_x, sizeof(double) * e->n, cudaMemcpyHostToDevice, e->stream));
    if (qlaunch(e, (const void*)k_qcp_solve_nspace, e->ctx, e->ctx.mu, (const double*)(host_warm_x ? e->ctx.hb : nullptr), rtol, max_iter))
        return -1;
    CK(cudaMemcpyAsync(host_vec, e->ctx.mu, sizeof(double) * mn, cudaMemcpyDeviceToHost, e->stream));
    return qread_sc(e, sc);
}

int abipgpu_qcp_get_vec(abipgpu_qcp* e, int id, double* host, long len) {
    CK(cudaSetDevice(e->device));
    const double* src = id == 0 ? e->u : id == 1 ? e->v : id == 2 ? e->ut : id == 3 ? e->r : nullptr;
    if (!src || len > e->l) return -1;
    CK(cudaMemcpyAsync(host, src, sizeof(double) * len, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

int abipgpu_qcp_set_vec(abipgpu_qcp* e, int id, const double* host, long len) {
    CK(cudaSetDevice(e->device));
    double* dst = id == 0 ? e->u : id == 1 ? e->v : nullptr;
    if (!dst || len > e->l) return -1;
    CK(cudaMemcpyAsync(dst, host, sizeof(double) * len, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}

double abipgpu_qcp_a_coef(const abipgpu_qcp* e) { return e->ctx.a_coef; }

void abipgpu_qcp_counters(const abipgpu_qcp* e, long* n_iter, long* n_cg, long* n_inner, double* kernel_ms) {
    *n_iter = e->n_iter; *n_cg = e->n_cg; *n_inner = e->n_inner; *kernel_ms = e->kernel_ms;
}

}  // extern "C"
