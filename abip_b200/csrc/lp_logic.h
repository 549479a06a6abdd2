// lp_logic.h -- scalar control logic of the ADMM loop shared by the host solver (lp_host.cpp) and the device-resident
// loops of the batch kernel (lp_engine.cu: k_batch, kinds BATCH_INNER / BATCH_BBSEARCH).  Plain functions of the
// 64-double scalar block an ADMM iteration / BB round reduces; compiled for both sides so that a problem advanced on
// the device takes exactly the branches the host loop would take.
#pragma once
#include <math.h>
#include "../../include/abip_gpu.h"

#if defined(__CUDACC__)
#define ABIP_HD __host__ __device__ inline
#else
#define ABIP_HD inline
#endif

struct LpResid {  // struct ABIP_RESIDUALS, include/abip.h:178-196 (the part the loop control reads)
    double res_pri, res_dual, rel_gap, res_infeas, res_unbdd, ct_x_by_tau, bt_y_by_tau, tau, kap;
};

struct LpResidIn {  // per-solve constants of calc_residuals
    double nm_b, nm_c, sc_b, sc_c, scale;
    int normalize;
};

ABIP_HD double lp_safediv_pos(double x, double y) { return y < 1e-18 ? x / 1e-18 : x / y; }  // SAFEDIV_POS, glbopts.h:157-158

// Q-norm criterion from one 13-scalar group (src/abip.c:1972-1992)
ABIP_HD double lp_qnorm_value(const double* g) {
    const double S_PR = g[0], BTY = g[3], UU_Y = g[4], S_DR = g[5], CTX = g[8], UU_X = g[9], VV = g[10];
    const double tau = g[11], kap = g[12];
    const double gap = BTY - CTX - kap;
    const double Q = S_PR + S_DR + gap * gap;
    const double nrm = 1 + sqrt((UU_Y + UU_X + tau * tau) + (VV + kap * kap));
    return sqrt(Q) / nrm;
}

// iterate_Q_norm_resd decision (abip.c:2040-2050): returns the criterion, sets *avg_criterion
ABIP_HD double lp_qnorm_decide(const double* sc, double max_admm_iters, int* avg_criterion) {
    const double q_cur = lp_qnorm_value(sc + ABIPGPU_SC_S_PR);
    const double q_avg = sc[ABIPGPU_SC_HAS_AVG] != 0 ? lp_qnorm_value(sc + ABIPGPU_SC_AVG_BASE) : sqrt(max_admm_iters) / 1.0;
    if (q_avg < q_cur) { *avg_criterion = 1; return q_avg; }
    *avg_criterion = 0;
    return q_cur;
}

// calc_residuals (src/abip.c:458-535) evaluated from the sums the ADMM kernel already reduced
ABIP_HD void lp_calc_residuals(const LpResidIn& in, const double* sc, int avg_criterion, LpResid* r) {
    const double* g = sc + (avg_criterion ? ABIPGPU_SC_AVG_BASE : ABIPGPU_SC_S_PR);
    const double W_AX = g[1], W_PR = g[2], BTY = g[3], W_ATYS = g[6], W_DR = g[7], CTX = g[8];
    const bool nz = in.normalize != 0;
    const double nrm = nz ? (in.scale * in.sc_c * in.sc_b) : 1.0;
    const double sb = nz ? (in.sc_b * in.scale) : 1.0, scn = nz ? (in.sc_c * in.scale) : 1.0;
    r->tau = fabs(g[11]);
    r->kap = fabs(g[12]) / nrm;
    const double nmpr_tau = sqrt(W_PR) / sb, nm_A_x_tau = sqrt(W_AX) / sb;
    const double nmdr_tau = sqrt(W_DR) / scn, nm_At_ys_tau = sqrt(W_ATYS) / scn;
    r->bt_y_by_tau = BTY / nrm;
    r->ct_x_by_tau = CTX / nrm;
    r->res_infeas = r->bt_y_by_tau > 0 ? in.nm_b * nm_At_ys_tau / r->bt_y_by_tau : NAN;
    r->res_unbdd = r->ct_x_by_tau < 0 ? in.nm_c * nm_A_x_tau / -r->ct_x_by_tau : NAN;
    const double bt_y = lp_safediv_pos(r->bt_y_by_tau, r->tau), ct_x = lp_safediv_pos(r->ct_x_by_tau, r->tau);
    r->res_pri = lp_safediv_pos(nmpr_tau / (1 + in.nm_b), r->tau);
    r->res_dual = lp_safediv_pos(nmdr_tau / (1 + in.nm_c), r->tau);
    r->rel_gap = fabs(ct_x - bt_y) / (1 + fabs(ct_x) + fabs(bt_y));
}

ABIP_HD int lp_has_converged(double eps, int pfeasopt, const LpResid* r, long ipm_iter, long admm_iter) {  // abip.c:1613-1641
    if (r->res_pri < eps && (r->res_dual < eps || pfeasopt) && r->rel_gap < eps) return ABIP_SOLVED;
    if (r->res_unbdd < eps && ipm_iter > 0 && admm_iter > 0) return ABIP_UNBOUNDED;
    if (r->res_infeas < eps && ipm_iter > 0 && admm_iter > 0) return ABIP_INFEASIBLE;
    return 0;
}

// One safeguarded Barzilai-Borwein step (src/adaptive.c:170-247) from the five inner products of a lookback round.
// Returns the action: 0 = stop the search (beta holds the result), 1 = continue with carry 1 (beta_prev <- beta,
// v_prev = [v_y; (mu/beta)/u_x]), 2 = continue with carry 2 (state hand-over only).
ABIP_HD int lp_bb_step(const double* sc, double eps_cor, double eps_pen, double* beta_prev, double* beta_out) {
    const double utut = sc[ABIPGPU_SC_BB_UTUT], utv = sc[ABIPGPU_SC_BB_UTV], uu = sc[ABIPGPU_SC_BB_UU],
                 vv = sc[ABIPGPU_SC_BB_VV], uv = sc[ABIPGPU_SC_BB_UV];
    const double norm_ut = sqrt(utut), norm_u = sqrt(uu), norm_v = sqrt(vv);
    const double alpha_SD = vv / utv, alpha_MG = utv / utut, gamma_SD = vv / uv, gamma_MG = uv / uu;
    const double alpha_ss = (2 * alpha_MG > alpha_SD) ? alpha_MG : alpha_SD - 0.5 * alpha_MG;
    const double gamma_ss = (2 * gamma_MG > gamma_SD) ? gamma_MG : gamma_SD - 0.5 * gamma_MG;
    const double alpha_cor = utv / (norm_v * norm_ut), gamma_cor = uv / (norm_v * norm_u);
    double beta;
    if (alpha_cor > eps_cor && gamma_cor > eps_cor) beta = sqrt(alpha_ss * gamma_ss);
    else if (alpha_cor > eps_cor && gamma_cor <= eps_cor) beta = alpha_ss;
    else if (alpha_cor <= eps_cor && gamma_cor > eps_cor) beta = gamma_ss;
    else beta = *beta_prev;
    const double diff = fabs(beta - *beta_prev);
    if (diff > 0 && diff <= eps_pen) {
        *beta_out = (beta + *beta_prev) / 2;
        return 0;
    }
    *beta_out = beta;
    if (diff > eps_pen) {
        *beta_prev = beta;
        return 1;
    }
    return 2;
}

// Device-resident inner ADMM loop of one outer iteration (src/abip.c:2131-2214) -- arguments and exit codes
struct LpInnerArgs {
    long j0, k0, j_end, cap;  // first inner index, global ADMM counter, inner_stopper, iterations per launch at most
    long max_admm_iters, max_ipm_iters, ipm_iter, restart_thresh;
    double mu, beta, gamma, eps;
    int final_check, pfeasopt, half_update, avg_in;
    LpResidIn rin;
};
enum {
    LP_INNER_CONTINUE = 0,   // launch cap reached: call again from j0 + iterations done
    LP_INNER_CONVERGED = 1,  // criterion < gamma * mu: the inner loop is over (abip.c:2173-2188)
    LP_INNER_FINISHED = 2,   // final_check: converged / iteration limit -> the solve is over (abip.c:2190-2211)
    LP_INNER_STOPPER = 3,    // j reached inner_stopper
    LP_INNER_HOST = 4,       // restart bookkeeping ahead (k >= restart_thresh): the host steps from here on
    // device-resident OUTER loop (LpSolveArgs, k_batch kind BATCH_SOLVE) in addition:
    LP_SOLVE_DONE = 5,       // the check after an inner loop found convergence / the ADMM limit (abip.c:2225-2249): get_solution
    LP_SOLVE_IPM = 6,        // max_ipm_iters outer iterations done (abip.c:2296)
    LP_SOLVE_FAIL = 7        // the LOQO mu rule met min(u_i v_i) <= 0 (the reference asserts, abip.c:962-965)
};

// ---- mu rules (src/abip.c:753-992, selection :2251-2277), shared by the host loop and the device-resident outer loop ----
struct LpMuState {
    double mu, sigma, gamma, dynamic_sigma;
    int final_check, double_check;
};
struct LpMuParams {
    double eps, sp, sparsity_ratio, dynamic_sigma_second, dynamic_x, hybrid_thresh;
    int hybrid_mu;
    long n_plus_1;
};
// which rule update_mu applies: 0 none, 1 the table rule (update_barrier), 2 update_barrier_dynamic_2, 3 the LOQO rule
// (update_barrier_dynamic: needs min and sum of u_i v_i over the (x, tau) tail)
ABIP_HD int lp_mu_rule(LpMuState* st, const LpMuParams& p) {
    if (p.hybrid_mu) {
        if (p.dynamic_sigma_second > 0.0 && st->mu < p.hybrid_thresh * p.eps) {
            st->dynamic_sigma = p.dynamic_sigma_second;
            return 3;
        } else if (p.dynamic_sigma_second == 0.0 && st->mu < p.hybrid_thresh * p.eps) {
            st->dynamic_sigma = p.dynamic_sigma_second;
            return 1;
        } else if (st->dynamic_sigma < 0.0) {
            return 2;
        }
        return 0;
    }
    if (st->dynamic_sigma == 0.0) return 1;
    if (st->dynamic_sigma < 0.0) return 2;
    return 3;
}
// table-driven mu rule (src/abip.c:753-921)
ABIP_HD void lp_update_barrier(LpMuState* st, const LpMuParams& p, const LpResid& r) {
    const double ratio = st->mu / p.eps;
    const double err = fmax(fmax(r.res_pri, r.res_dual), r.rel_gap) / p.eps;
    const bool dense = fmax(p.sp, p.sparsity_ratio) > 0.4 || fmin(p.sp, p.sparsity_ratio) > 0.1;
    const double lo[8] = {10.0, 1.0, 0.5, 0.1, 0.05, 0.01, 0.005, 0.001};
    const double gam[8] = {dense ? 2.0 : 3.0, 1.0, 0.9, 0.8, 0.7, 0.6, 0.5, 0.4};
    double gamma = 0.3, sigma = st->sigma;
    for (int q = 0; q < 8; ++q)
        if (ratio > lo[q]) { gamma = gam[q]; break; }
    if (dense) {
        if (err > 6 && err <= 10) sigma = 0.5;
        else if (err > 3 && err <= 6) { sigma = 0.6; gamma *= 0.8; }
        else if (err > 1 && err <= 3) { st->final_check = 1; gamma *= 0.4; sigma = ratio < 0.1 ? 0.8 : 0.7; }
    } else {
        if (err > 6 && err <= 10) { sigma = 0.82; gamma *= 0.8; }
        else if (err > 4 && err <= 6) { sigma = 0.84; gamma *= 0.6; }
        else if (err > 3 && err <= 4) { sigma = 0.85; gamma *= 0.5; st->final_check = 1; }
        else if (err > 1 && err <= 3) {
            st->final_check = 1;
            if (ratio < 0.1) {
                if (st->double_check) { sigma = 0.9; gamma *= 0.4; st->double_check = 0; }
                else { sigma = 1.0; gamma *= 0.1; st->double_check = 1; }
            } else { sigma = 0.88; gamma *= 0.4; }
        }
    }
    st->mu *= sigma;
    st->sigma = sigma;
    st->gamma = gamma;
}
// src/abip.c:982-992; eta = dynamic_sigma (parity trap 6)
ABIP_HD void lp_update_barrier_dynamic_2(LpMuState* st, const LpMuParams& p) {
    st->mu *= fmin(p.dynamic_x * st->mu, pow(st->mu, st->dynamic_sigma));
}
// LOQO rule (src/abip.c:930-977) from min / sum of u_i v_i; -1: invalid (min <= 0)
ABIP_HD int lp_update_barrier_dynamic(LpMuState* st, const LpMuParams& p, double minxs, double sumxs) {
    if (!(minxs > 0.0)) return -1;
    const double xs = sumxs / (double)p.n_plus_1;
    const double ksi = minxs / xs;
    double sigma = fmin(0.05 * (1 - ksi) / ksi, 2.0);
    sigma = fmax(0.1 * sigma * sigma * sigma, st->dynamic_sigma);
    st->mu *= sigma;
    return 0;
}
// inner_stopper of an outer iteration (src/abip.c:2098-2112)
ABIP_HD long lp_inner_stopper(double spmin, double mu, long max_admm_iters) {
    if (spmin > 0.5) return (long)round(pow(mu, -0.35));
    if (spmin > 0.2) return (long)round(pow(mu, -1.0));
    return max_admm_iters;
}

// Device-resident OUTER loop (src/abip.c:2093-2295): inner loops, the convergence check after each of them, the mu rule,
// re-initialisation and the Barzilai-Borwein search run inside ONE batched step until the solve ends, the launch cap
// (in.cap ADMM iterations) is reached or the restart bookkeeping needs the host.  in.ipm_iter / in.j0 / in.k0 / in.mu /
// in.beta / in.gamma / in.final_check / in.avg_in carry the state in; the state comes back in the scalar block
// (ABIPGPU_SC_LOOP_*).
struct LpSolveArgs {
    LpInnerArgs in;
    LpMuParams mp;
    double sigma, dynamic_sigma;
    int double_check;
    int resume_inner;  // 1: the launch starts inside the inner loop of outer iteration in.ipm_iter at j0 (prologue done)
    int adaptive, adaptive_lookback;
    double eps_cor, eps_pen;
};
