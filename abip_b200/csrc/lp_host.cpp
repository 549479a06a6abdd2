// lp_host.cpp -- host side of the ABIP-LP engine: the reference's linsys plugin symbols (group (1) of
// include/abip_gpu.h) and the outer IPM / inner ADMM / Barzilai-Borwein control flow of the solver entry
// (group (2)).  Only scalar decisions happen here; every vector operation is a CUDA kernel in lp_engine.cu.
// Each function cites the reference lines whose behaviour it reproduces (paths under src/abip-lp/).
#include "lp_engine.h"
#include "order_host.h"
#include "lp_logic.h"
#include <thread>

#include <algorithm>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace {

constexpr double kMinScale = 1e-3, kMaxScale = 1e3;  // linsys/common.c:4-5, src/normalize.c:5-6
constexpr double kEpsTol = 1e-18;                    // include/glbopts.h:157
constexpr double kIndeterminateTol = 1e-9;           // include/glbopts.h:161

inline double safediv_pos(double x, double y) { return y < kEpsTol ? x / kEpsTol : x / y; }  // glbopts.h:158

double now_ms() {  // src/util.c:73-102 (CLOCK_MONOTONIC, milliseconds)
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec / 1e6;
}

// ---- SIGINT polling, same contract as src/ctrlc.c:62-93 -------------------------------------------------
// The handler is process-global but solves may run concurrently (abip_gpu_batch_main: one host thread per problem in
// flight): the listener is reference-counted under a mutex -- only the FIRST solve saves the host's handler and clears
// the flag, only the LAST one restores it, so an interrupt delivered during a batch reaches every solve and the host
// process (e.g. Python's KeyboardInterrupt handler) gets its own handler back afterwards.
volatile sig_atomic_t g_interrupted = 0;
struct sigaction g_old_action;
std::mutex g_sig_mu;
int g_sig_users = 0;
void on_sigint(int) { g_interrupted = 1; }
void start_interrupt_listener() {
    std::lock_guard<std::mutex> lk(g_sig_mu);
    if (g_sig_users++ > 0) return;
    struct sigaction act;
    g_interrupted = 0;
    act.sa_flags = 0;
    sigemptyset(&act.sa_mask);
    act.sa_handler = on_sigint;
    sigaction(SIGINT, &act, &g_old_action);
}
void end_interrupt_listener() {
    std::lock_guard<std::mutex> lk(g_sig_mu);
    if (g_sig_users <= 0 || --g_sig_users > 0) return;
    struct sigaction cur;
    sigaction(SIGINT, &g_old_action, &cur);
}

double norm2(const double* a, long n) {
    double s = 0;
    for (long i = 0; i < n; ++i) s += a[i] * a[i];
    return std::sqrt(s);
}

}  // namespace

// =========================================================================================================
// (1) linsys plugin
// =========================================================================================================
struct ABIP_LIN_SYS_WORK {
    abipgpu_lp* eng;
    abip_int tot_cg_its;        // linsys/indirect.h:26-28
    abip_float total_solve_time;
    int failed;                 // latched error of accum_by_A / accum_by_Atrans (their signature has no error channel)
};

extern "C" {

abip_int abip_copy_A_matrix(ABIPMatrix** dstp, const ABIPMatrix* src) {  // common.c:10-41 (1 = ok, 0 = fail)
    const abip_int nnz = src->p[src->n];
    ABIPMatrix* A = (ABIPMatrix*)calloc(1, sizeof(ABIPMatrix));
    if (!A) return 0;
    A->m = src->m;
    A->n = src->n;
    A->x = (abip_float*)malloc(sizeof(abip_float) * nnz);
    A->i = (abip_int*)malloc(sizeof(abip_int) * nnz);
    A->p = (abip_int*)malloc(sizeof(abip_int) * (src->n + 1));
    if (!A->x || !A->i || !A->p) {
        free(A->x); free(A->i); free(A->p); free(A);
        return 0;
    }
    std::copy(src->x, src->x + nnz, A->x);
    std::copy(src->i, src->i + nnz, A->i);
    std::copy(src->p, src->p + src->n + 1, A->p);
    *dstp = A;
    return 1;
}

void abip_free_A_matrix(ABIPMatrix* A) {  // common.c:100-118
    if (!A) return;
    free(A->x); free(A->i); free(A->p); free(A);
}

abip_int abip_validate_lin_sys(const ABIPMatrix* A) {  // common.c:44-96
    if (!A->x || !A->i || !A->p) {
        printf("ERROR: incomplete data!\n");
        return -1;
    }
    for (abip_int j = 0; j < A->n; ++j) {
        if (A->p[j] == A->p[j + 1]) printf("WARN: the %li-th column empty!\n", (long)j);
        else if (A->p[j] > A->p[j + 1]) {
            printf("ERROR: the column pointers decreases!\n");
            return -1;
        }
    }
    const abip_int nnz = A->p[A->n];
    if (((abip_float)nnz / A->m > A->n) || nnz <= 0) {
        printf("ERROR: the number of nonzeros in A = %li, outside of valid range!\n", (long)nnz);
        return -1;
    }
    abip_int rmax = 0;
    for (abip_int k = 0; k < nnz; ++k) rmax = std::max(rmax, A->i[k]);
    if (rmax > A->m - 1) {
        printf("ERROR: the number of rows in A is inconsistent with input dimension!\n");
        return -1;
    }
    return 0;
}

// Equilibration of A (common.c:150-565).  One generic sweep -- "scale columns by f(column), then rows by
// f(row)" -- instantiated for the pc (sqrt of 1-norm), origin (2-norm), ruiz (sqrt of inf-norm, repeated) and qp
// (sqrt(min*max)) variants; D and E accumulate the products (:524-532).
void abip_normalize_A(ABIPMatrix* A, const ABIPSettings* stgs, ABIPScaling* scal) {
    const abip_int m = A->m, n = A->n, nnz = A->p[n];
    double* D = (double*)malloc(sizeof(double) * m);
    double* E = (double*)malloc(sizeof(double) * n);
    std::fill(D, D + m, 1.0);
    std::fill(E, E + n, 1.0);
    const double min_row = kMinScale * std::sqrt((double)n), max_row = kMaxScale * std::sqrt((double)n);
    const double min_col = kMinScale * std::sqrt((double)m), max_col = kMaxScale * std::sqrt((double)m);
    std::vector<double> dt(m), dt2(m);
    enum Kind { PC, ORIGIN, RUIZ, QP };
    auto clamp = [](double v, double lo, double hi) { return v < lo ? 1.0 : (v > hi ? hi : v); };
    auto sweep = [&](Kind kind) {
        for (abip_int j = 0; j < n; ++j) {  // column factor, applied immediately
            double acc = 0, mn = 0;
            const abip_int c1 = A->p[j], c2 = A->p[j + 1];
            for (abip_int k = c1; k < c2; ++k) {
                const double a = std::fabs(A->x[k]);
                if (kind == PC) acc += a;
                else if (kind == ORIGIN) acc += a * a;
                else acc = std::max(acc, a);
            }
            double e;
            if (kind == QP) {  // sqrt(min nonzero) * sqrt(max), :423-428
                mn = acc;
                for (abip_int k = c1; k < c2; ++k) {
                    const double a = std::fabs(A->x[k]);
                    if (a <= mn && a > 0) mn = a;
                }
                e = std::sqrt(mn) * std::sqrt(acc);
            } else {
                e = std::sqrt(acc);
            }
            e = clamp(e, min_col, max_col);
            const double inv = 1.0 / e;
            for (abip_int k = c1; k < c2; ++k) A->x[k] *= inv;
            E[j] *= e;
        }
        std::fill(dt.begin(), dt.end(), 0.0);
        for (abip_int k = 0; k < nnz; ++k) {  // row statistic
            const double a = std::fabs(A->x[k]);
            double& d = dt[A->i[k]];
            if (kind == PC) d += a;
            else if (kind == ORIGIN) d += a * a;
            else if (a >= d) d = a;
        }
        if (kind == QP) {
            dt2 = dt;
            for (abip_int k = 0; k < nnz; ++k) {
                const double a = std::fabs(A->x[k]);
                if (a <= dt2[A->i[k]] && a > 0) dt2[A->i[k]] = a;
            }
        }
        for (abip_int i = 0; i < m; ++i) {
            const double v = (kind == QP) ? std::sqrt(dt[i] * dt2[i]) : std::sqrt(dt[i]);
            dt[i] = clamp(v, min_row, max_row);
            D[i] *= dt[i];
        }
        for (abip_int k = 0; k < nnz; ++k) A->x[k] /= dt[A->i[k]];
    };
    const bool say = stgs->verbose != 0;  // the reference prints these unconditionally
    if (stgs->pc_ruiz_rescale) { sweep(PC); if (say) printf("Done the pc rescaling!\n"); }
    if (stgs->origin_rescale) { sweep(ORIGIN); if (say) printf("Done the origin rescaling!\n"); }
    if (stgs->pc_ruiz_rescale) {
        for (abip_int it = 0; it < stgs->ruiz_iter; ++it) sweep(RUIZ);
        if (say) printf("Done the ruiz rescaling!\n");
    }
    if (stgs->qp_rescale) { sweep(QP); if (say) printf("Done the QP rescaling\n"); }
    // mean row / column 2-norms of the scaled matrix (:535-557)
    std::fill(dt.begin(), dt.end(), 0.0);
    for (abip_int k = 0; k < nnz; ++k) dt[A->i[k]] += A->x[k] * A->x[k];
    scal->mean_norm_row_A = 0.0;
    for (abip_int i = 0; i < m; ++i) scal->mean_norm_row_A += std::sqrt(dt[i]) / m;
    scal->mean_norm_col_A = 0.0;
    for (abip_int j = 0; j < n; ++j)
        scal->mean_norm_col_A += norm2(A->x + A->p[j], A->p[j + 1] - A->p[j]) / n;
    if (stgs->scale != 1)
        for (abip_int k = 0; k < nnz; ++k) A->x[k] *= stgs->scale;
    scal->D = D;
    scal->E = E;
}

void abip_un_normalize_A(ABIPMatrix* A, const ABIPSettings* stgs, const ABIPScaling* scal) {  // common.c:570-594
    for (abip_int j = 0; j < A->n; ++j)
        for (abip_int k = A->p[j]; k < A->p[j + 1]; ++k) A->x[k] *= scal->D[A->i[k]] * (scal->E[j] / stgs->scale);
}

char* abip_get_lin_sys_method(const ABIPMatrix* A, const ABIPSettings* stgs) {  // indirect.c:8-18
    char* str = (char*)malloc(128);
    snprintf(str, 128, "sparse-indirect (B200 CUDA PCG), nnz in A = %li, CG tol ~ 1/iter^(%2.2f)", (long)A->p[A->n],
             stgs->cg_rate);
    return str;
}

char* abip_get_lin_sys_summary(ABIPLinSysWork* p, const ABIPInfo* info) {  // indirect.c:20-33
    char* str = (char*)malloc(128);
    snprintf(str, 128, "\tLin-sys: avg # CG iterations: %2.2f, avg solve time: %1.2es\n",
             (abip_float)p->tot_cg_its / (info->admm_iter + 1), p->total_solve_time / (info->admm_iter + 1) / 1e3);
    p->tot_cg_its = 0;
    p->total_solve_time = 0;
    return str;
}

ABIPLinSysWork* abip_init_lin_sys_work(const ABIPMatrix* A, const ABIPSettings* stgs) {  // indirect.c:282-318
    ABIPLinSysWork* p = (ABIPLinSysWork*)calloc(1, sizeof(ABIPLinSysWork));
    if (!p) return nullptr;
    const char* dev = getenv("ABIP_GPU_DEVICE");
    p->eng = abipgpu_lp_create(A->m, A->n, A->p, A->i, A->x, stgs, dev ? atoi(dev) : 0);
    if (!p->eng) {
        free(p);
        return nullptr;
    }
    return p;
}

void abip_free_lin_sys_work(ABIPLinSysWork* p) {  // indirect.c:141-203
    if (!p) return;
    abipgpu_lp_destroy(p->eng);
    free(p);
}

abip_int abip_solve_lin_sys(const ABIPMatrix* A, const ABIPSettings* stgs, ABIPLinSysWork* p, abip_float* b,
                            const abip_float* s, abip_int iter) {  // indirect.c:393-434
    (void)A;
    (void)stgs;
    const double t0 = now_ms();
    int its = 0;
    if (p->failed) return -1;  // an earlier accum_by_A / accum_by_Atrans failed (no error channel there): fail here
    if (abipgpu_lp_solve_host(p->eng, b, s, (long)iter, &its) != 0) return -1;
    if (iter >= 0) p->tot_cg_its += its;
    p->total_solve_time += now_ms() - t0;
    return 0;
}

void abip_accum_by_Atrans(const ABIPMatrix* A, ABIPLinSysWork* p, const abip_float* x, abip_float* y) {
    (void)A;
    // the reference signature has no error channel: latch the error, the next solve_lin_sys returns -1 (=> ABIP_FAILED in
    // the caller, src/abip.c:2137-2140) instead of taking the host process (MATLAB) down
    if (abipgpu_lp_spmv_host(p->eng, 1, x, y, 1) != 0) {
        fprintf(stderr, "[abip_gpu] accum_by_Atrans failed\n");
        p->failed = 1;
    }
}

void abip_accum_by_A(const ABIPMatrix* A, ABIPLinSysWork* p, const abip_float* x, abip_float* y) {
    (void)A;
    if (abipgpu_lp_spmv_host(p->eng, 0, x, y, 1) != 0) {
        fprintf(stderr, "[abip_gpu] accum_by_A failed\n");
        p->failed = 1;
    }
}

}  // extern "C"

// =========================================================================================================
// (2) solver entry
// =========================================================================================================
struct Resid {  // struct ABIP_RESIDUALS, include/abip.h:178-196
    abip_int last_ipm_iter = -1, last_admm_iter = -1;
    double res_pri = NAN, res_dual = NAN, rel_gap = NAN, res_infeas = NAN, res_unbdd = NAN;
    double ct_x_by_tau = NAN, bt_y_by_tau = NAN, tau = NAN, kap = NAN;
};

struct ABIP_GPU_WORK {  // device-resident replacement of struct ABIP_WORK (include/abip.h:126-176)
    abip_int m = 0, n = 0;    // global dimensions
    abip_int c0 = 0, nl = 0;  // multi-GPU: this rank owns columns [c0, c0 + nl) (single GPU: all of them)
    int dist_G = 1, dist_rank = 0;
    ABIPMatrix* A = nullptr;  // scaled private copy (COPYAMATRIX behaviour, abip.c:1799-1810)
    ABIPScaling scal{nullptr, nullptr, 0, 0};
    bool have_scal = false;
    ABIPSettings stgs;  // private copy: the reference mutates the caller's struct (avg_criterion, dynamic_sigma,
                        // max_admm_iters; SURVEY.md parity trap 5) -- we mutate this copy instead
    ABIPSettings stgs0; // as passed to init; every solve starts from it, so repeated solves are independent
    double sp = 0;
    abipgpu_lp* eng = nullptr;
    std::vector<double> b, c;  // scaled
    bool device_loops = false;      // batch engines: inner ADMM loop and BB search run on the device (lp_engine.cu: k_batch)
    std::vector<int> rperm, cperm;  // multi-GPU: locality ordering applied to the host copy of A (new -> old; empty: identity)
    double sigma = 0, gamma = 0, mu = 1, beta = 1;
    int final_check = 0, double_check = 0;
    double sc_b = 1, sc_c = 1, nm_b = 0, nm_c = 0;
    double sc[ABIPGPU_SC_COUNT];  // last scalar block of an ADMM iteration
    abip_int tot_cg_its = 0;
    double total_solve_ms = 0, total_adapt_ms = 0;
    ABIPGpuStats last_stats;
    FILE* trace = nullptr;
};

namespace {

int validate(const ABIPData* d) {  // src/abip.c:1646-1734
    const ABIPSettings* s = d->stgs;
    if (d->m <= 0 || d->n <= 0) {
        printf("m and n must both be greater than 0; m = %li, n = %li\n", (long)d->m, (long)d->n);
        return -1;
    }
    if (d->m > d->n) {
        printf("WARN: m larger than n, problem likely degenerate\n");
        return -1;
    }
    if (abip_validate_lin_sys(d->A) < 0) {
        printf("invalid linear system input data\n");
        return -1;
    }
    struct { bool bad; const char* msg; } checks[] = {
        {s->max_ipm_iters <= 0, "max_ipm_iters must be positive"},
        {s->max_admm_iters <= 0, "max_admm_iters must be positive"},
        {s->eps <= 0, "eps tolerance must be positive"},
        {s->alpha <= 0 || s->alpha >= 2, "alpha must be in (0,2)"},
        {s->rho_y <= 0, "rho_y must be positive (1e-3 works well)."},
        {s->scale <= 0, "scale must be positive (1 works well)."},
        {s->eps_cor <= 0, "eps_cor tolerance must be positive."},
        {s->eps_pen <= 0, "eps_pen tolerance must be positive."},
        {s->adaptive_lookback <= 0, "adaptive_lookback must be positive."},
        {s->hybrid_mu > 0 && s->dynamic_sigma >= 0, "when use hybrid mu strategy, dynamic_sigma must be negative."},
    };
    for (auto& ck : checks)
        if (ck.bad) {
            printf("%s\n", ck.msg);
            return -1;
        }
    return 0;
}

void fill_nan(double* a, abip_int n) {
    for (abip_int i = 0; i < n; ++i) a[i] = NAN;
}

abip_int failure(abip_int m, abip_int n, ABIPSolution* sol, ABIPInfo* info, abip_int status, const char* msg,
                 const char* ststr) {  // src/abip.c:219-303
    if (info) {
        info->res_pri = info->res_dual = info->rel_gap = info->res_infeas = info->res_unbdd = NAN;
        info->pobj = info->dobj = NAN;
        info->ipm_iter = info->admm_iter = -1;
        info->status_val = status;
        info->solve_time = NAN;
        snprintf(info->status, sizeof(info->status), "%s", ststr);
    }
    if (sol) {
        if (n > 0) {
            if (!sol->x) sol->x = (double*)malloc(sizeof(double) * n);
            if (!sol->s) sol->s = (double*)malloc(sizeof(double) * n);
            fill_nan(sol->x, n);
            fill_nan(sol->s, n);
        }
        if (m > 0) {
            if (!sol->y) sol->y = (double*)malloc(sizeof(double) * m);
            fill_nan(sol->y, m);
        }
    }
    printf("Failure:%s\n", msg);
    return status;  // (abip_gpu_solve releases the SIGINT listener in its epilogue guard)
}

static LpResidIn resid_in(const ABIP_GPU_WORK* w) {
    return LpResidIn{w->nm_b, w->nm_c, w->sc_b, w->sc_c, w->stgs.scale, (int)w->stgs.normalize};
}

// calc_residuals (src/abip.c:458-535) evaluated from the sums the ADMM kernel already reduced (lp_logic.h)
void calc_residuals(ABIP_GPU_WORK* w, Resid* r, abip_int ipm_iter, abip_int admm_iter) {
    if (admm_iter && r->last_admm_iter == admm_iter) return;
    r->last_ipm_iter = ipm_iter;
    r->last_admm_iter = admm_iter;
    LpResid q;
    lp_calc_residuals(resid_in(w), w->sc, (int)w->stgs.avg_criterion, &q);
    r->res_pri = q.res_pri; r->res_dual = q.res_dual; r->rel_gap = q.rel_gap; r->res_infeas = q.res_infeas;
    r->res_unbdd = q.res_unbdd; r->ct_x_by_tau = q.ct_x_by_tau; r->bt_y_by_tau = q.bt_y_by_tau; r->tau = q.tau; r->kap = q.kap;
}

abip_int has_converged(const ABIP_GPU_WORK* w, const Resid* r, abip_int ipm_iter, abip_int admm_iter) {  // :1613-1641
    LpResid q{r->res_pri, r->res_dual, r->rel_gap, r->res_infeas, r->res_unbdd, r->ct_x_by_tau, r->bt_y_by_tau, r->tau, r->kap};
    return lp_has_converged(w->stgs.eps, (int)w->stgs.pfeasopt, &q, (long)ipm_iter, (long)admm_iter);
}

// mu rules (src/abip.c:753-992) and their selection (:2251-2277): the scalar logic lives in lp_logic.h (shared with the
// device-resident outer loop); min / sum of u_i v_i for the LOQO rule are reduced on the device
static LpMuParams mu_params(const ABIP_GPU_WORK* w) {
    const ABIPSettings& s = w->stgs;
    return LpMuParams{s.eps, w->sp, s.sparsity_ratio, s.dynamic_sigma_second, s.dynamic_x, s.hybrid_thresh, (int)s.hybrid_mu,
                      (long)w->n + 1};
}
int update_mu(ABIP_GPU_WORK* w, const Resid* r) {
    ABIPSettings& s = w->stgs;
    LpMuState st{w->mu, w->sigma, w->gamma, s.dynamic_sigma, w->final_check, w->double_check};
    const LpMuParams mp = mu_params(w);
    int rc = 0;
    const int rule = lp_mu_rule(&st, mp);
    if (rule == 1) {
        const LpResid q{r->res_pri, r->res_dual, r->rel_gap, r->res_infeas, r->res_unbdd, r->ct_x_by_tau, r->bt_y_by_tau, r->tau, r->kap};
        lp_update_barrier(&st, mp, q);
    } else if (rule == 2) {
        lp_update_barrier_dynamic_2(&st, mp);
    } else if (rule == 3) {
        double sc[ABIPGPU_SC_COUNT];
        if (abipgpu_lp_mu_stats(w->eng, (int)s.avg_criterion, sc) != 0) return -1;
        if (lp_update_barrier_dynamic(&st, mp, sc[ABIPGPU_SC_MIN_XS], sc[ABIPGPU_SC_SUM_XS]) < 0) {  // the reference assert(0)s here (:962-965)
            printf("Invalid xisi < 0 \n");
            rc = -1;
        }
    }
    w->mu = st.mu; w->sigma = st.sigma; w->gamma = st.gamma; s.dynamic_sigma = st.dynamic_sigma;
    w->final_check = st.final_check; w->double_check = st.double_check;
    return rc;
}

// Barzilai-Borwein search for beta (src/adaptive.c:34-256): the two ADMM steps and the five inner products of a
// lookback round are one kernel launch; the safeguarded spectral step below is scalar work.
int adaptive_search(ABIP_GPU_WORK* w, abip_int iter) {
    const ABIPSettings& s = w->stgs;
    if (s.adaptive_lookback <= 0) return -1;
    const double t0 = now_ms();
    if (abipgpu_lp_bb_begin(w->eng) != 0) return -1;
    if (w->device_loops && !w->trace) {  // batch engines: the whole search is one batched step (lp_engine.cu: BATCH_BBSEARCH)
        double scb[ABIPGPU_SC_COUNT];
        if (abipgpu_lp_bb_search(w->eng, iter, w->mu, (int)s.adaptive_lookback, s.eps_cor, s.eps_pen, scb) != 0) return -1;
        w->tot_cg_its += (abip_int)scb[ABIPGPU_SC_LOOP_CG];
        w->beta = scb[ABIPGPU_SC_LOOP_BETA];
        w->total_adapt_ms += now_ms() - t0;
        return 0;
    }
    double beta_prev = 1.0, beta = 0.0;
    int carry = 0;
    double sc[ABIPGPU_SC_COUNT];
    for (abip_int i = 0; i < s.adaptive_lookback; ++i) {
        if (abipgpu_lp_bb_round(w->eng, carry, iter, w->mu, beta_prev, sc) != 0) return -1;
        w->tot_cg_its += (abip_int)(sc[ABIPGPU_SC_CG_ITS] + sc[ABIPGPU_SC_CG_ITS2]);
        const double bp_before = beta_prev;
        const int action = lp_bb_step(sc, s.eps_cor, s.eps_pen, &beta_prev, &beta);
        if (w->trace)
            fprintf(w->trace, "bbround %ld carry %d beta_prev %.17g beta %.17g dots %.10e %.10e %.10e %.10e %.10e cg %d %d\n",
                    (long)i, carry, bp_before, beta, sc[ABIPGPU_SC_BB_UTUT], sc[ABIPGPU_SC_BB_UTV], sc[ABIPGPU_SC_BB_UU],
                    sc[ABIPGPU_SC_BB_VV], sc[ABIPGPU_SC_BB_UV], (int)sc[ABIPGPU_SC_CG_ITS], (int)sc[ABIPGPU_SC_CG_ITS2]);
        if (action == 0) break;
        carry = action;  // 1: u_prev = u; v_prev = [v_y; (mu/beta)/u_x] (:230-242); 2: u_prev = u; v_prev = v (:243-247)
    }
    w->beta = beta;
    w->total_adapt_ms += now_ms() - t0;
    return 0;
}

void print_header_line(const ABIP_GPU_WORK* w) {
    static const char* cols[] = {" ipm iter ", " admm iter ", "     mu ", " pri res ", " dua res ", " rel gap ",
                                 " pri obj ", " dua obj ", " kap/tau ", " time (s)"};
    (void)w;
    for (int i = 0; i < 150; ++i) printf("-");
    printf("\n");
    for (int i = 0; i < 10; ++i) printf("%s%s", cols[i], i < 9 ? "|" : "\n");
    for (int i = 0; i < 150; ++i) printf("=");
    printf("\n");
}

void print_summary(const ABIP_GPU_WORK* w, abip_int i, abip_int k, const Resid* r, double t0) {  // abip.c:1418-1463
    printf("%*i|", 10, (int)i);
    printf("%*i|", 11, (int)k);
    printf("%*.2e|", 8, w->mu);
    printf("%*.2e|%*.2e|%*.2e|", 9, r->res_pri, 9, r->res_dual, 9, r->rel_gap);
    printf("%*.2e|%*.2e|", 9, safediv_pos(r->ct_x_by_tau, r->tau), 9, safediv_pos(r->bt_y_by_tau, r->tau));
    printf("%*.2e|%*.2e\n", 9, safediv_pos(r->kap, r->tau), 9, (now_ms() - t0) / 1e3);
    fflush(stdout);
}

// get_solution + get_info (src/abip.c:1308-1414), un_normalize_sol (src/normalize.c:133-158)
int get_solution(ABIP_GPU_WORK* w, ABIPSolution* sol, ABIPInfo* info, Resid* r, abip_int ipm_iter, abip_int admm_iter) {
    const abip_int m = w->m, n = w->n, nl = w->nl, l = m + nl + 1;  // l: length of this rank's (u, v)
    calc_residuals(w, r, ipm_iter, admm_iter);
    if (!sol->x) sol->x = (double*)malloc(sizeof(double) * n);
    if (!sol->y) sol->y = (double*)malloc(sizeof(double) * m);
    if (!sol->s) sol->s = (double*)malloc(sizeof(double) * n);
    std::vector<double> uu(l), vv(l);
    const int avg = (int)w->stgs.avg_criterion;
    if (abipgpu_lp_get_vec(w->eng, avg ? ABIPGPU_VEC_UAVGC : ABIPGPU_VEC_U, uu.data(), l) != 0 ||
        abipgpu_lp_get_vec(w->eng, avg ? ABIPGPU_VEC_VAVGC : ABIPGPU_VEC_V, vv.data(), l) != 0)
        return -1;
    // multi-GPU: each rank returns its own column shard of x and s (zeros elsewhere; the caller sums the shards)
    std::fill(sol->x, sol->x + n, 0.0);
    std::fill(sol->s, sol->s + n, 0.0);
    if (w->cperm.empty()) {
        std::copy(uu.begin(), uu.begin() + m, sol->y);
        std::copy(uu.begin() + m, uu.begin() + m + nl, sol->x + w->c0);
        std::copy(vv.begin() + m, vv.begin() + m + nl, sol->s + w->c0);
    } else {  // engine order -> caller's order
        for (abip_int i = 0; i < m; ++i) sol->y[w->rperm[i]] = uu[i];
        for (abip_int j = 0; j < nl; ++j) {
            sol->x[w->cperm[w->c0 + j]] = uu[m + j];
            sol->s[w->cperm[w->c0 + j]] = vv[m + j];
        }
    }
    enum { SOLVED, INDET, INFEAS, UNBDD } kind;
    const abip_int sv = info->status_val;
    if (sv == ABIP_UNFINISHED) {
        if (r->tau > kIndeterminateTol && r->tau > r->kap) kind = SOLVED;
        else if (norm2(uu.data(), l) < kIndeterminateTol * std::sqrt((double)l)) kind = INDET;
        else if (-r->bt_y_by_tau < r->ct_x_by_tau) kind = INFEAS;
        else kind = UNBDD;
    } else if (sv == ABIP_SOLVED || sv == ABIP_SOLVED_INACCURATE) kind = SOLVED;
    else if (sv == ABIP_INFEASIBLE || sv == ABIP_INFEASIBLE_INACCURATE) kind = INFEAS;
    else kind = UNBDD;
    const bool inacc = (sv == 0);
    auto scale_all = [](double* a, abip_int len, double f) { for (abip_int i = 0; i < len; ++i) a[i] *= f; };
    if (kind == SOLVED) {
        const double f = safediv_pos(1.0, r->tau);
        scale_all(sol->x, n, f); scale_all(sol->y, m, f); scale_all(sol->s, n, f);
        info->status_val = inacc ? ABIP_SOLVED_INACCURATE : ABIP_SOLVED;
        snprintf(info->status, sizeof(info->status), "%s", inacc ? "Solved/Inaccurate" : "Solved");
    } else if (kind == INDET) {
        fill_nan(sol->x, n); fill_nan(sol->y, m); fill_nan(sol->s, n);
        info->status_val = ABIP_INDETERMINATE;
        snprintf(info->status, sizeof(info->status), "Indeterminate");
    } else if (kind == INFEAS) {
        scale_all(sol->y, m, 1 / r->bt_y_by_tau); scale_all(sol->s, n, 1 / r->bt_y_by_tau);
        fill_nan(sol->x, n);
        info->status_val = inacc ? ABIP_INFEASIBLE_INACCURATE : ABIP_INFEASIBLE;
        snprintf(info->status, sizeof(info->status), "%s", inacc ? "Infeasible/Inaccurate" : "Infeasible");
    } else {
        scale_all(sol->x, n, -1 / r->ct_x_by_tau);
        fill_nan(sol->y, m); fill_nan(sol->s, n);
        info->status_val = inacc ? ABIP_UNBOUNDED_INACCURATE : ABIP_UNBOUNDED;
        snprintf(info->status, sizeof(info->status), "%s", inacc ? "Unbounded/Inaccurate" : "Unbounded");
    }
    if (w->stgs.normalize) {
        const double* D = w->scal.D; const double* E = w->scal.E;
        for (abip_int i = 0; i < n; ++i) sol->x[i] /= (E[i] * w->sc_b);
        for (abip_int i = 0; i < m; ++i) sol->y[i] /= (D[i] * w->sc_c);
        for (abip_int i = 0; i < n; ++i) sol->s[i] *= E[i] / (w->sc_c * w->stgs.scale);
    }
    info->ipm_iter = ipm_iter + 1;
    info->admm_iter = admm_iter + 1;
    info->res_infeas = r->res_infeas;
    info->res_unbdd = r->res_unbdd;
    if (kind == SOLVED) {
        info->rel_gap = r->rel_gap; info->res_pri = r->res_pri; info->res_dual = r->res_dual;
        info->pobj = r->ct_x_by_tau / r->tau; info->dobj = r->bt_y_by_tau / r->tau;
    } else if (kind == UNBDD) {
        info->rel_gap = info->res_pri = info->res_dual = NAN;
        info->pobj = info->dobj = -INFINITY;
    } else if (kind == INFEAS) {
        info->rel_gap = info->res_pri = info->res_dual = NAN;
        info->pobj = info->dobj = INFINITY;
    }
    return 0;
}

void print_footer(ABIP_GPU_WORK* w, const ABIPInfo* info) {  // src/abip.c:1465-1596 (abridged)
    for (int i = 0; i < 150; ++i) printf("-");
    printf("\nStatus: %s\n", info->status);
    printf("Timing: Solve time: %1.2es\n", info->solve_time / 1e3);
    printf("\tLin-sys: avg # CG iterations: %2.2f\n", (double)w->tot_cg_its / (info->admm_iter + 1));
    printf("\tBarzilai-Borwein spectral method: avg step time: %1.2es\n",
           w->total_adapt_ms / (info->admm_iter + 1) / 1e3);
    printf("Error metrics:\nprimal res = %.4e, dual res = %.4e, rel gap = %.4e\n", info->res_pri, info->res_dual,
           info->rel_gap);
    printf("c'x = %.4f, b'y = %.4f\n", info->pobj, info->dobj);
    for (int i = 0; i < 150; ++i) printf("=");
    printf("\n");
}

}  // namespace

extern "C" {

void abip_gpu_set_default_settings(ABIPData* d) {  // src/util.c:288-329, mexfile/abip_mex.c:320-341
    ABIPSettings* s = d->stgs;
    s->max_ipm_iters = 500; s->max_admm_iters = 1000000; s->eps = 1e-3; s->alpha = 1.8; s->cg_rate = 2.0;
    s->normalize = 1; s->scale = 1.0; s->rho_y = 1e-3; s->sparsity_ratio = 0.01;
    s->adaptive = 1; s->eps_cor = 0.2; s->eps_pen = 0.1; s->adaptive_lookback = 20;
    s->dynamic_x = 0.8; s->dynamic_eta = 1.1;
    s->restart_fre = 1000; s->restart_thresh = 100000;
    s->origin_rescale = 0; s->pc_ruiz_rescale = 1; s->qp_rescale = 0; s->ruiz_iter = 10;
    s->hybrid_mu = 1; s->dynamic_sigma = -1.0; s->hybrid_thresh = 1000; s->dynamic_sigma_second = 0.5;
    s->half_update = 0; s->avg_criterion = 0;
    s->verbose = 1; s->warm_start = 0;
    s->max_time = 3600; s->pfeasopt = 0;
}

static ABIPGpuWork* gpu_init_impl(const ABIPData* d, ABIPInfo* info, int rank, int G);

// contiguous column blocks balanced by nonzeros: rank r owns columns [c0, c0 + nl)
void abip_gpu_column_partition(abip_int n, const abip_int* Ap, abip_int world, abip_int rank, abip_int* c0, abip_int* nl) {
    const abip_int nnz = Ap[n];
    auto cut = [&](abip_int r) {
        if (r >= world) return n;
        const double target = (double)nnz * (double)r / (double)world;
        abip_int lo = 0, hi = n;
        while (lo < hi) {
            const abip_int mid = (lo + hi) / 2;
            if ((double)Ap[mid] < target) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    *c0 = cut(rank);
    *nl = cut(rank + 1) - *c0;
}

ABIPGpuWork* abip_gpu_init(const ABIPData* d, ABIPInfo* info) { return gpu_init_impl(d, info, 0, 1); }

// Multi-GPU (one process per GPU): every rank passes the FULL problem; the equilibration is computed redundantly on
// every rank (identical D, E), then rank r keeps a contiguous block of columns balanced by nonzeros.  The caller
// must then exchange the IPC handles: abip_gpu_comm_export on every rank, all-gather, abip_gpu_comm_connect.
ABIPGpuWork* abip_gpu_init_dist(const ABIPData* d, ABIPInfo* info, abip_int rank, abip_int world) {
    if (world < 1 || world > 8 || rank < 0 || rank >= world) return nullptr;
    return gpu_init_impl(d, info, (int)rank, (int)world);
}

abip_int abip_gpu_comm_export(ABIPGpuWork* w, void* handle64) { return abipgpu_lp_comm_export(w->eng, handle64); }

abip_int abip_gpu_comm_connect(ABIPGpuWork* w, const void* handles) {
    return abipgpu_lp_comm_connect(w->eng, w->dist_G, w->dist_rank, handles);
}

void abip_gpu_partition(const ABIPGpuWork* w, abip_int* c0, abip_int* nl) {
    *c0 = w->c0;
    *nl = w->nl;
}

// Multi-GPU: the locality ordering (order_host.h) is applied ONCE to the host copy of the (scaled) matrix, identically on
// every rank (a pure function of the structure), before the column blocks are cut: the m-space is replicated and
// exchanged by row index, so all ranks must agree on the row order; the engines then keep the order they are given.
static void reorder_host_matrix(ABIPGpuWork* w) {
    const abip_int m = w->m, n = w->n;
    const long nnz = w->A->p[n];
    if (nnz >= 2147483647L) return;
    const int threads = std::max(1, std::min(8, (int)std::thread::hardware_concurrency()));
    auto par = [&](long cnt, auto fn) { parallel_for(cnt, threads, fn); };
    std::vector<int> at_ptr(n + 1), at_idx(nnz), a_ptr(m + 1, 0), a_idx(nnz);
    for (abip_int j = 0; j <= n; ++j) at_ptr[j] = (int)w->A->p[j];
    for (long k = 0; k < nnz; ++k) { at_idx[k] = (int)w->A->i[k]; a_ptr[w->A->i[k] + 1]++; }
    for (abip_int i = 0; i < m; ++i) a_ptr[i + 1] += a_ptr[i];
    {
        std::vector<int> fill(a_ptr.begin(), a_ptr.end() - 1);
        for (abip_int j = 0; j < n; ++j)
            for (long k = w->A->p[j]; k < w->A->p[j + 1]; ++k) a_idx[fill[w->A->i[k]]++] = (int)j;
    }
    std::vector<int> rn2o, cn2o;
    sjds::locality_order((int)m, (int)n, a_ptr, a_idx, at_ptr, at_idx, sjds::kLongRow, &rn2o, &cn2o, par);
    bool ident = true;
    for (abip_int i = 0; i < m && ident; ++i) ident = rn2o[i] == i;
    for (abip_int j = 0; j < n && ident; ++j) ident = cn2o[j] == j;
    if (ident) return;
    std::vector<int> ro2n(m);
    for (abip_int i = 0; i < m; ++i) ro2n[rn2o[i]] = (int)i;
    abip_int* np_ = (abip_int*)malloc(sizeof(abip_int) * (n + 1));
    abip_int* ni = (abip_int*)malloc(sizeof(abip_int) * nnz);
    double* nx = (double*)malloc(sizeof(double) * nnz);
    np_[0] = 0;
    for (abip_int j = 0; j < n; ++j) np_[j + 1] = np_[j] + (w->A->p[cn2o[j] + 1] - w->A->p[cn2o[j]]);
    par(n, [&](long j0, long j1, int) {
        for (long j = j0; j < j1; ++j) {
            long q = np_[j];
            for (long k = w->A->p[cn2o[j]]; k < w->A->p[cn2o[j] + 1]; ++k, ++q) {
                ni[q] = ro2n[w->A->i[k]];
                nx[q] = w->A->x[k];
            }
        }
    });
    free(w->A->p); free(w->A->i); free(w->A->x);
    w->A->p = np_; w->A->i = ni; w->A->x = nx;
    w->rperm = std::move(rn2o);
    w->cperm = std::move(cn2o);
}

static ABIPGpuWork* gpu_init_impl(const ABIPData* d, ABIPInfo* info, int rank, int G) {  // ABIP(init) + init_work, abip.c:1739-1841, 2341-2389
    if (!d || !info) {
        printf("ERROR: Missing ABIPData or ABIPInfo input\n");
        return nullptr;
    }
    if (validate(d) < 0) {
        printf("ERROR: Validation returned failure\n");
        return nullptr;
    }
    const double t0 = now_ms();
    ABIPGpuWork* w = new ABIPGpuWork();
    w->m = d->m;
    w->n = d->n;
    w->stgs = *d->stgs;
    w->stgs0 = *d->stgs;
    w->sp = d->sp;
    if (d->stgs->verbose) {
        char* meth = abip_get_lin_sys_method(d->A, d->stgs);
        for (int i = 0; i < 150; ++i) printf("-");
        printf("\n\tABIP-B200 - First-Order Interior-Point Solver, device-resident ADMM engine (sm_100a)\n");
        for (int i = 0; i < 150; ++i) printf("-");
        printf("\nLin-sys: %s\n", meth);
        free(meth);
        printf("eps = %.2e, alpha = %.2f, max_ipm_iters = %i, max_admm_iters = %i, normalize = %i\n"
               "scale = %2.2f, adaptive = %i, adaptive_lookback = %i, rho_y = %.2e\n",
               d->stgs->eps, d->stgs->alpha, (int)d->stgs->max_ipm_iters, (int)d->stgs->max_admm_iters,
               (int)d->stgs->normalize, d->stgs->scale, (int)d->stgs->adaptive, (int)d->stgs->adaptive_lookback,
               d->stgs->rho_y);
        printf("Variables n = %i, constraints m = %i\n", (int)d->n, (int)d->m);
    }
    const char* dev = getenv("ABIP_GPU_DEVICE");
    w->dist_G = G;
    w->dist_rank = rank;
    w->c0 = 0;
    w->nl = w->n;
    // batches of small LPs equilibrate on the host: the device version costs ~100 driver calls per problem (40 tiny
    // launches), which serialise on the context lock and were 2/3 of the wall time of a batch; on the host it is a
    // fraction of a millisecond per problem and runs in parallel over the worker threads (results are bit-identical)
    const bool host_scaling = G > 1 || abipgpu_batch_attached() || getenv("ABIP_GPU_HOST_SCALING") != nullptr;
    if (!host_scaling) {
        // single GPU: the caller's matrix goes to the device as it is and is equilibrated there (bit-identical to
        // abip_normalize_A); the host keeps only D and E
        if (w->stgs.normalize) {
            w->scal.D = (double*)malloc(sizeof(double) * w->m);
            w->scal.E = (double*)malloc(sizeof(double) * w->n);
            w->have_scal = true;
            w->eng = abipgpu_lp_create_scaling(w->m, w->n, d->A->p, d->A->i, d->A->x, &w->stgs, dev ? atoi(dev) : 0,
                                               w->scal.D, w->scal.E, &w->scal.mean_norm_row_A, &w->scal.mean_norm_col_A);
        } else {
            w->eng = abipgpu_lp_create(w->m, w->n, d->A->p, d->A->i, d->A->x, &w->stgs, dev ? atoi(dev) : 0);
        }
        if (!w->eng) {
            printf("ERROR: init_lin_sys_work failure\n");
            abip_gpu_finish(w);
            return nullptr;
        }
        abipgpu_lp_set_global_n(w->eng, w->n);
        if (d->stgs->verbose) {
            char buf[1280];
            abipgpu_lp_describe(w->eng, buf, sizeof(buf));
            printf("Engine: %s\n", buf);
        }
        info->setup_time = now_ms() - t0;
        if (d->stgs->verbose) printf("Setup time: %1.2es\n", info->setup_time / 1e3);
        return w;
    }
    if (!abip_copy_A_matrix(&w->A, d->A)) {
        printf("ERROR: copy A matrix failed\n");
        delete w;
        return nullptr;
    }
    if (w->stgs.normalize) {
        // multi-GPU: every rank equilibrates the FULL matrix on its own GPU (bit-identical to abip_normalize_A, which took
        // ~25 s of host time per rank at cfg4)
        bool done = false;
        if (G > 1 && !getenv("ABIP_GPU_HOST_SCALING")) {
            w->scal.D = (double*)malloc(sizeof(double) * w->m);
            w->scal.E = (double*)malloc(sizeof(double) * w->n);
            done = abipgpu_equilibrate(w->m, w->n, w->A->p, w->A->i, w->A->x, &w->stgs, dev ? atoi(dev) : 0, w->scal.D, w->scal.E,
                                       &w->scal.mean_norm_row_A, &w->scal.mean_norm_col_A) == 0;
            if (!done) {
                free(w->scal.D); free(w->scal.E);
                w->scal.D = w->scal.E = nullptr;
                abip_free_A_matrix(w->A);
                w->A = nullptr;
                if (!abip_copy_A_matrix(&w->A, d->A)) { delete w; return nullptr; }
            }
        }
        if (!done) abip_normalize_A(w->A, &w->stgs, &w->scal);
        w->have_scal = true;
    }
    if (G > 1 && !getenv("ABIP_GPU_NO_REORDER")) reorder_host_matrix(w);
    if (G > 1) {  // contiguous column blocks balanced by nonzeros
        if (w->stgs.half_update) {
            printf("ERROR: half_update is not supported by the multi-GPU engine\n");
            abip_gpu_finish(w);
            return nullptr;
        }
        abip_gpu_column_partition(w->n, w->A->p, G, rank, &w->c0, &w->nl);
        if (w->nl <= 0) {
            printf("ERROR: empty column block on rank %d\n", rank);
            abip_gpu_finish(w);
            return nullptr;
        }
    }
    {
        const abip_int p0 = w->A->p[w->c0];
        std::vector<abip_int> lp(w->nl + 1);
        for (abip_int j = 0; j <= w->nl; ++j) lp[j] = w->A->p[w->c0 + j] - p0;
        // the m-space of a sharded engine is replicated and exchanged by row index: every rank must keep the caller's order
        if (G > 1) abipgpu_lp_request_order(0);
        w->eng = abipgpu_lp_create(w->m, w->nl, lp.data(), w->A->i + p0, w->A->x + p0, &w->stgs, dev ? atoi(dev) : 0);
        if (G > 1) abipgpu_lp_request_order(1);
    }
    if (!w->eng) {
        printf("ERROR: init_lin_sys_work failure\n");
        abip_gpu_finish(w);
        return nullptr;
    }
    abipgpu_lp_set_global_n(w->eng, w->n);
    if (G > 1) {
        // the Jacobi preconditioner M = 1/diag(AA') (indirect.c:36-79) involves ALL columns: replace the engine's
        // local one by the global diagonal so that the replicated PCG is identical on every rank
        std::vector<double> M(w->m, 0.0);
        const abip_int nnz = w->A->p[w->n];
        for (abip_int k = 0; k < nnz; ++k) M[w->A->i[k]] += w->A->x[k] * w->A->x[k];
        for (abip_int i = 0; i < w->m; ++i) M[i] = 1.0 / M[i];
        if (abipgpu_lp_set_vec(w->eng, ABIPGPU_VEC_M, M.data(), w->m) != 0) {
            abip_gpu_finish(w);
            return nullptr;
        }
    }
    if (d->stgs->verbose) {
        char buf[1280];
        abipgpu_lp_describe(w->eng, buf, sizeof(buf));
        printf("Engine: %s\n", buf);
    }
    info->setup_time = now_ms() - t0;
    if (d->stgs->verbose) printf("Setup time: %1.2es\n", info->setup_time / 1e3);
    return w;
}

abip_int abip_gpu_solve(ABIPGpuWork* w, const ABIPData* d, ABIPSolution* sol, ABIPInfo* info) {  // abip.c:2056-2297
    if (!d || !sol || !info || !w || !d->b || !d->c) {
        printf("ERROR: ABIP_NULL input\n");
        return ABIP_FAILED;
    }
    const abip_int m = w->m, n = w->n;
    struct SolvingGuard {  // lock-step batches: the executor launches when every thread inside a solve is waiting
        abipgpu_lp* e;
        explicit SolvingGuard(abipgpu_lp* e_) : e(e_) { abipgpu_lp_batch_solving(e, +1); }
        ~SolvingGuard() { abipgpu_lp_batch_solving(e, -1); }
    } solving_guard(w->eng);
    w->stgs = w->stgs0;
    ABIPSettings& s = w->stgs;
    const double t0 = now_ms();
    const double max_time = s.max_time;
    start_interrupt_listener();
    info->status_val = ABIP_UNFINISHED;
    Resid r;
    ABIPGpuStats* st = abipgpu_lp_stats(w->eng);
    memset(st, 0, sizeof(*st));
    abipgpu_lp_solve_timer(w->eng, 0);
    // ONE epilogue for every way out of the solve (normal returns and the failure() paths): the solve timer is stopped,
    // the statistics are copied, deferred batch operations are dropped and the SIGINT listener is released
    struct SolveEpilogue {
        ABIP_GPU_WORK* w;
        ABIPGpuStats* st;
        ~SolveEpilogue() {
            abipgpu_lp_solve_timer(w->eng, 1);
            w->last_stats = *st;
            abipgpu_lp_drop_pending(w->eng);
            end_interrupt_listener();
        }
    } solve_epilogue{w, st};
    w->tot_cg_its = 0;
    w->total_adapt_ms = 0;
    w->device_loops = abipgpu_lp_is_batch(w->eng) != 0 && getenv("ABIP_GPU_BATCH_HOST_LOOPS") == nullptr;

    // ---- update_work (abip.c:1843-1927) ----
    w->nm_b = norm2(d->b, m);
    w->nm_c = norm2(d->c, n);
    w->b.assign(d->b, d->b + m);
    w->c.assign(d->c, d->c + n);
    w->sc_b = w->sc_c = 1.0;
    if (s.normalize) {  // normalize_b_c, src/normalize.c:11-40
        for (abip_int i = 0; i < n; ++i) w->c[i] /= w->scal.E[i];
        w->sc_c = w->scal.mean_norm_row_A / std::max(norm2(w->c.data(), n), kMinScale);
        for (abip_int i = 0; i < m; ++i) w->b[i] /= w->scal.D[i];
        w->sc_b = w->scal.mean_norm_col_A / std::max(norm2(w->b.data(), m), kMinScale);
        for (abip_int i = 0; i < n; ++i) w->c[i] *= w->sc_c * s.scale;
        for (abip_int i = 0; i < m; ++i) w->b[i] *= w->sc_b * s.scale;
    }
    const double spmin = std::min(w->sp, s.sparsity_ratio), spmax = std::max(w->sp, s.sparsity_ratio);
    if (spmax > 0.4 || (spmin > 0.1 && spmin < 0.2)) { w->sigma = 0.3; w->gamma = 2.0; }
    else if (spmin > 0.2) { w->sigma = 0.5; w->gamma = 3.0; }
    else { w->sigma = 0.8; w->gamma = 3.0; }
    w->final_check = 0;
    w->double_check = 0;
    w->mu = 1.0;
    w->beta = 1.0;
    if (s.warm_start) {
        // The reference's warm_start_vars (abip.c:307-357) overwrites every non-NaN entry with sqrt(mu/beta)
        // unless built with NOVALIDATE, i.e. it degenerates to the cold start.
        printf("WARN: warm_start behaves as in the reference build: iterates restart from sqrt(mu/beta)\n");
    }
    {
        const double *pb = w->b.data(), *pc = w->c.data() + w->c0;
        const double *pD = s.normalize ? w->scal.D : nullptr, *pE = s.normalize ? w->scal.E + w->c0 : nullptr;
        std::vector<double> qb, qc, qD, qE;
        if (!w->cperm.empty()) {  // multi-GPU with locality ordering: the engine's matrix is P_r A P_c
            qb.resize(m); qc.resize(w->nl);
            for (abip_int i = 0; i < m; ++i) qb[i] = w->b[w->rperm[i]];
            for (abip_int j = 0; j < w->nl; ++j) qc[j] = w->c[w->cperm[w->c0 + j]];
            pb = qb.data(); pc = qc.data();
            if (s.normalize) {
                qD.resize(m); qE.resize(w->nl);
                for (abip_int i = 0; i < m; ++i) qD[i] = w->scal.D[w->rperm[i]];
                for (abip_int j = 0; j < w->nl; ++j) qE[j] = w->scal.E[w->cperm[w->c0 + j]];
                pD = qD.data(); pE = qE.data();
            }
        }
        if (abipgpu_lp_cold_start(w->eng, w->mu, w->beta) != 0 || abipgpu_lp_set_problem(w->eng, pb, pc, pD, pE) != 0)
            return failure(m, n, sol, info, ABIP_FAILED, "error in update_work", "Failure");
    }

    if (s.verbose) print_header_line(w);

    FILE* trace = nullptr;  // ABIP_GPU_TRACE=<file>: one line per ADMM iteration / BB search (debug + parity tests)
    if (const char* tf = getenv("ABIP_GPU_TRACE")) trace = fopen(tf, "w");
    w->trace = trace;
    struct TraceCloser { FILE* f; ABIP_GPU_WORK* w; ~TraceCloser() { if (f) fclose(f); w->trace = nullptr; } } trace_closer{trace, w};

    abip_int k = 0;
    abip_int i_start = 0, j_start = 0;
    bool resume = false;  // the host loop takes over inside the inner loop of outer iteration i_start at j_start
    if (w->device_loops && !trace && !s.verbose && getenv("ABIP_GPU_BATCH_HOST_OUTER") == nullptr) {
        // Batch engines: the OUTER loop runs on the device as well (lp_engine.cu: k_batch, BATCH_SOLVE; decisions in
        // lp_logic.h) -- a problem is one batched step (one launch per `cap` ADMM iterations) instead of ~45 host round trips.
        // The host is back in charge when the solve is over, between launches (time limit, SIGINT), or for the restart
        // bookkeeping (k >= restart_thresh), where it continues with single steps from the state handed back.
        LpSolveArgs S;
        memset(&S, 0, sizeof(S));
        S.in.cap = 4096;
        S.in.max_ipm_iters = s.max_ipm_iters;
        S.in.restart_thresh = s.restart_thresh;
        S.in.eps = s.eps;
        S.in.pfeasopt = (int)s.pfeasopt; S.in.half_update = (int)s.half_update;
        S.in.rin = resid_in(w);
        S.adaptive = (int)s.adaptive; S.adaptive_lookback = (int)s.adaptive_lookback;
        S.eps_cor = s.eps_cor; S.eps_pen = s.eps_pen;
        abip_int i = 0, j = 0;
        int code = LP_INNER_CONTINUE;
        for (;;) {
            S.in.j0 = j; S.in.k0 = k; S.in.ipm_iter = i; S.in.max_admm_iters = s.max_admm_iters;
            S.in.mu = w->mu; S.in.beta = w->beta; S.in.gamma = w->gamma;
            S.in.final_check = w->final_check; S.in.avg_in = (int)s.avg_criterion;
            S.mp = mu_params(w);
            S.sigma = w->sigma; S.dynamic_sigma = s.dynamic_sigma; S.double_check = w->double_check;
            S.resume_inner = resume ? 1 : 0;
            if (abipgpu_lp_solve_loop(w->eng, &S, w->sc) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in project_lin_sys", "Failure");
            code = (int)w->sc[ABIPGPU_SC_LOOP_EXIT];
            k += (abip_int)w->sc[ABIPGPU_SC_LOOP_ITERS];
            w->tot_cg_its += (abip_int)w->sc[ABIPGPU_SC_LOOP_CG];
            i = (abip_int)w->sc[ABIPGPU_SC_LOOP_I];
            j = (abip_int)w->sc[ABIPGPU_SC_LOOP_J];
            s.avg_criterion = (abip_int)w->sc[ABIPGPU_SC_LOOP_AVG];
            w->mu = w->sc[ABIPGPU_SC_LOOP_MU]; w->beta = w->sc[ABIPGPU_SC_LOOP_BETA];
            w->sigma = w->sc[ABIPGPU_SC_LOOP_SIGMA]; w->gamma = w->sc[ABIPGPU_SC_LOOP_GAMMA];
            s.dynamic_sigma = w->sc[ABIPGPU_SC_LOOP_DYN];
            const int flags = (int)w->sc[ABIPGPU_SC_LOOP_FLAGS];
            w->final_check = flags & 1; w->double_check = (flags >> 1) & 1;
            if (g_interrupted) return failure(m, n, sol, info, ABIP_SIGINT, "Interrupted", "Interrupted");
            if (code != LP_INNER_CONTINUE) break;
            resume = true;  // launch cap reached inside an inner loop: go on from (i, j, k)
            if ((now_ms() - t0) / 1e3 > max_time) {  // wall clock, checked between launches
                printf("Timelimit reached. \n");
                s.max_admm_iters = (abip_int)(k * 1.05);
            }
        }
        if (code == LP_SOLVE_FAIL) return failure(m, n, sol, info, ABIP_FAILED, "error in mu update", "Failure");
        if (code == LP_INNER_FINISHED || code == LP_SOLVE_DONE) {
            calc_residuals(w, &r, i, k);
            info->status_val = has_converged(w, &r, i, k);
            if (get_solution(w, sol, info, &r, i, k) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in get_solution", "Failure");
            info->solve_time = now_ms() - t0;
            return info->status_val;
        }
        if (code == LP_SOLVE_IPM) {
            i_start = s.max_ipm_iters;  // skip the host loop: straight to the final get_solution
        } else {  // LP_INNER_HOST: restart bookkeeping ahead
            w->device_loops = false;
            i_start = i;
            j_start = j;
            resume = true;
        }
    }
    for (abip_int i = i_start; i < s.max_ipm_iters; ++i) {  // outer loop
        const bool resumed = resume && i == i_start;
        abip_int inner_stopper;
        if (spmin > 0.5) inner_stopper = (abip_int)std::round(std::pow(w->mu, -0.35));
        else if (spmin > 0.2) inner_stopper = (abip_int)std::round(std::pow(w->mu, -1));
        else inner_stopper = s.max_admm_iters;
        if (!resumed && abipgpu_lp_outer_prologue(w->eng, (int)s.avg_criterion) != 0)
            return failure(m, n, sol, info, ABIP_FAILED, "error in outer prologue", "Failure");

        // a solve is over inside the inner loop when final_check finds convergence or an iteration limit (abip.c:2190-2211)
        auto finish_in_loop = [&](abip_int kk) -> abip_int {
            if (s.verbose && kk > 0) print_summary(w, i, kk, &r, t0);
            if (get_solution(w, sol, info, &r, i, kk) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in get_solution", "Failure");
            info->solve_time = now_ms() - t0;
            if (s.verbose) print_footer(w, info);
            return info->status_val;
        };
        for (abip_int j = resumed ? j_start : 0; j < inner_stopper;) {  // inner loop
            if (w->device_loops && !trace) {
                // batch engines: up to `cap` iterations per batched step, decisions taken on the device (lp_logic.h)
                LpInnerArgs L;
                L.j0 = j; L.k0 = k; L.j_end = inner_stopper; L.cap = 48;
                L.max_admm_iters = s.max_admm_iters; L.max_ipm_iters = s.max_ipm_iters; L.ipm_iter = i;
                L.restart_thresh = s.restart_thresh;
                L.mu = w->mu; L.beta = w->beta; L.gamma = w->gamma; L.eps = s.eps;
                L.final_check = w->final_check; L.pfeasopt = (int)s.pfeasopt; L.half_update = (int)s.half_update;
                L.avg_in = (int)s.avg_criterion;
                L.rin = resid_in(w);
                if (abipgpu_lp_inner_loop(w->eng, &L, w->sc) != 0)
                    return failure(m, n, sol, info, ABIP_FAILED, "error in project_lin_sys", "Failure");
                const abip_int done = (abip_int)w->sc[ABIPGPU_SC_LOOP_ITERS];
                const int code = (int)w->sc[ABIPGPU_SC_LOOP_EXIT];
                k += done;
                w->tot_cg_its += (abip_int)w->sc[ABIPGPU_SC_LOOP_CG];
                s.avg_criterion = (abip_int)w->sc[ABIPGPU_SC_LOOP_AVG];
                if (g_interrupted) return failure(m, n, sol, info, ABIP_SIGINT, "Interrupted", "Interrupted");
                if (code == LP_INNER_HOST) {  // restart bookkeeping ahead: single steps from here on
                    w->device_loops = false;
                    j += done;
                    continue;
                }
                if (code == LP_INNER_CONVERGED) break;
                if (code == LP_INNER_FINISHED) {
                    calc_residuals(w, &r, i, k);
                    info->status_val = has_converged(w, &r, i, k);
                    return finish_in_loop(k);
                }
                j += done;
                continue;  // LP_INNER_STOPPER ends the loop through its condition, LP_INNER_CONTINUE goes on
            }
            if (abipgpu_lp_admm_iter(w->eng, j, k, w->mu, w->beta, w->sc) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in project_lin_sys", "Failure");
            w->tot_cg_its += (abip_int)w->sc[ABIPGPU_SC_CG_ITS];
            if (g_interrupted) return failure(m, n, sol, info, ABIP_SIGINT, "Interrupted", "Interrupted");
            k += 1;
            // iterate_Q_norm_resd decision (abip.c:2040-2050)
            int avg_c = 0;
            const double q = lp_qnorm_decide(w->sc, (double)s.max_admm_iters, &avg_c);
            s.avg_criterion = avg_c;
            if (trace)
                fprintf(trace, "it %ld %ld %ld %.17g %.17g %d %.17g %d\n", (long)i, (long)j, (long)k, w->mu, w->beta,
                        (int)w->sc[ABIPGPU_SC_CG_ITS], q, (int)s.avg_criterion);
            if (q < w->gamma * w->mu) {
                if (s.half_update && abipgpu_lp_clamp_v(w->eng) != 0)  // abip.c:2175-2186
                    return failure(m, n, sol, info, ABIP_FAILED, "error in clamp", "Failure");
                break;
            }
            if (w->final_check) {
                calc_residuals(w, &r, i, k);
                info->status_val = has_converged(w, &r, i, k);
                if (info->status_val != 0 || k + 1 >= s.max_admm_iters || i + 1 >= s.max_ipm_iters) return finish_in_loop(k);
            }
            ++j;
        }
        if ((now_ms() - t0) / 1e3 > max_time) {  // wall clock; the reference uses clock() CPU time (trap 9)
            printf("Timelimit reached. \n");
            s.max_admm_iters = (abip_int)(k * 1.05);
        }
        if (w->mu < s.eps) w->final_check = 1;
        calc_residuals(w, &r, i, k);
        if (s.verbose) print_summary(w, i, k, &r, t0);
        info->status_val = has_converged(w, &r, i, k);
        if (info->status_val != 0 || k + 1 >= s.max_admm_iters) {
            if (get_solution(w, sol, info, &r, i, k) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in get_solution", "Failure");
            info->solve_time = now_ms() - t0;
            if (s.verbose) print_footer(w, info);
            return info->status_val;
        }
        if (update_mu(w, &r) != 0) return failure(m, n, sol, info, ABIP_FAILED, "error in mu update", "Failure");
        const int avg = (int)s.avg_criterion;
        if (abipgpu_lp_reinit(w->eng, 0, w->sigma, avg) != 0)
            return failure(m, n, sol, info, ABIP_FAILED, "error in reinitialize_vars", "Failure");
        if (s.adaptive) {
            if (abipgpu_lp_reinit(w->eng, 1, w->sigma, avg) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in reinitialize_vars", "Failure");
            w->beta = 1;
            if (adaptive_search(w, k) < 0) return failure(m, n, sol, info, ABIP_FAILED, "error in adaptive", "Failure");
            if (trace) fprintf(trace, "bb %ld %.17g %.17g %.17g\n", (long)i, w->mu, w->sigma, w->beta);
            if (abipgpu_lp_reinit(w->eng, 2, w->sigma, avg) != 0)
                return failure(m, n, sol, info, ABIP_FAILED, "error in reinitialize_vars", "Failure");
        }
    }
    // max_ipm_iters exhausted.  The reference returns here without filling sol/info (abip.c:2296); we report the
    // last iterate instead (status Solved/Inaccurate or the certificate get_solution selects).
    if (get_solution(w, sol, info, &r, s.max_ipm_iters - 1, k) != 0)
        return failure(m, n, sol, info, ABIP_FAILED, "error in get_solution", "Failure");
    info->solve_time = now_ms() - t0;
    return info->status_val;
}

void abip_gpu_finish(ABIPGpuWork* w) {  // abip.c:2301-2336
    if (!w) return;
    abipgpu_lp_destroy(w->eng);
    abip_free_A_matrix(w->A);
    free(w->scal.D);
    free(w->scal.E);
    delete w;
}

abip_int abip_gpu_main(const ABIPData* d, ABIPSolution* sol, ABIPInfo* info) {  // abip.c:2393-2422
    abip_int status;
    ABIPGpuWork* w = abip_gpu_init(d, info);
    if (w) {
        abip_gpu_solve(w, d, sol, info);
        status = info->status_val;
    } else {
        status = failure(d ? d->m : -1, d ? d->n : -1, sol, info, ABIP_FAILED, "could not initialize work", "Failure");
    }
    abip_gpu_finish(w);
    return status;
}

void abip_gpu_get_stats(const ABIPGpuWork* w, ABIPGpuStats* out) { *out = w->last_stats; }

// Batch of independent LPs on ONE GPU (BASELINE.json configs[4]; the reference equivalent is a loop of ABIP(main)
// calls, one process per core).  `concurrency` host threads pull problems from a shared counter.
//   ctas_per_problem >= 2: every problem gets its own engine with a persistent grid of that many CTAs on its own
//     stream (several cooperative kernels share the SMs); one launch + one synchronisation per problem and step.
//   ctas_per_problem <= 1 (lock-step mode): one CTA per problem; the blocking steps of all problems in flight are
//     launched together by the batch executor of lp_engine.cu (one k_batch launch per step for the whole batch).
abip_int abip_gpu_batch_main(const ABIPData* const* problems, ABIPSolution* sols, ABIPInfo* infos, abip_int count,
                             abip_int concurrency, abip_int ctas_per_problem) {
    if (!problems || !sols || !infos || count <= 0) return -1;
    if (concurrency < 1) concurrency = 1;
    if (concurrency > count) concurrency = count;
    std::atomic<long> next{0};
    std::atomic<long> failed{0};
    void* exec = nullptr;
    if (ctas_per_problem <= 1) {
        const char* dev = getenv("ABIP_GPU_DEVICE");
        exec = abipgpu_batch_begin(dev ? atoi(dev) : 0, (int)concurrency);
        if (!exec) return -1;
        // device memory of the engines in flight, estimated from the largest of the first problems: both matrices (values,
        // indices, plans) + ~30 vectors of length m + n + 1
        size_t per_engine = 0;
        for (abip_int i = 0; i < std::min<abip_int>(count, 16); ++i)
            if (problems[i] && problems[i]->A && problems[i]->A->p) {
                const size_t nnz = (size_t)problems[i]->A->p[problems[i]->n], l = (size_t)problems[i]->m + problems[i]->n + 1;
                per_engine = std::max(per_engine, 40 * nnz + 30 * 8 * l + (size_t)65536);
            }
        abipgpu_batch_reserve(exec, per_engine * (size_t)concurrency * 5 / 4, (int)concurrency + 8);
    }
    // lock-step mode: set-up and tear-down of a problem are ~150 driver calls under the per-context lock; letting all
    // threads fight for it at once starves the executor's own launch + synchronise, so only a few threads at a time
    // may be in set-up / tear-down
    const char* gate_env = getenv("ABIP_GPU_BATCH_SETUP_GATE");
    int gate_free = gate_env ? std::max(1, atoi(gate_env)) : 8;
    std::mutex gate_mu;
    std::condition_variable gate_cv;
    auto gate_enter = [&] {
        std::unique_lock<std::mutex> lk(gate_mu);
        gate_cv.wait(lk, [&] { return gate_free > 0; });
        --gate_free;
    };
    auto gate_leave = [&] {
        {
            std::lock_guard<std::mutex> lk(gate_mu);
            ++gate_free;
        }
        gate_cv.notify_one();
    };
    // tear-down is a handful of stream-ordered frees: outside the gate (12 - 24 ms per problem inside it, profiles/r02_batch.md)
    const bool finish_gated = getenv("ABIP_GPU_BATCH_FINISH_GATE") ? atoi(getenv("ABIP_GPU_BATCH_FINISH_GATE")) != 0 : false;
    std::mutex stat_mu;
    double stat_ms[4] = {0, 0, 0, 0};  // per problem: waiting at the set-up gate, init, solve, finish (incl. its gate)
    auto worker = [&]() {
        if (exec) abipgpu_batch_attach(exec);
        else abipgpu_lp_request_grid((int)ctas_per_problem);
        for (;;) {
            const long i = next.fetch_add(1);
            if (i >= count) break;
            abip_int st;
            if (exec) {
                const double t_a = now_ms();
                gate_enter();
                const double t_b = now_ms();
                ABIPGpuWork* w = abip_gpu_init(problems[i], &infos[i]);
                gate_leave();
                const double t_c = now_ms();
                if (w) {
                    abip_gpu_solve(w, problems[i], &sols[i], &infos[i]);
                    st = infos[i].status_val;
                } else {
                    st = failure(problems[i] ? problems[i]->m : -1, problems[i] ? problems[i]->n : -1, &sols[i], &infos[i],
                                 ABIP_FAILED, "could not initialize work", "Failure");
                }
                const double t_d = now_ms();
                if (finish_gated) gate_enter();
                abip_gpu_finish(w);
                if (finish_gated) gate_leave();
                const double t_e = now_ms();
                {
                    std::lock_guard<std::mutex> lk(stat_mu);
                    stat_ms[0] += t_b - t_a; stat_ms[1] += t_c - t_b; stat_ms[2] += t_d - t_c; stat_ms[3] += t_e - t_d;
                }
            } else {
                st = abip_gpu_main(problems[i], &sols[i], &infos[i]);
            }
            if (st == ABIP_FAILED) failed.fetch_add(1);
        }
        abipgpu_lp_request_grid(0);
        abipgpu_batch_attach(nullptr);
    };
    std::vector<std::thread> pool;
    for (abip_int t = 0; t < concurrency; ++t) pool.emplace_back(worker);
    for (auto& th : pool) th.join();
    if (exec) {
        long launches = 0, items = 0;
        abipgpu_batch_end(exec, &launches, &items);
        if (getenv("ABIP_GPU_BATCH_VERBOSE"))
            printf("[abip_gpu] batch: %ld problems, %ld batched launches, %.1f steps per launch; host thread per problem: %.2f ms at "
                   "the set-up gate, %.2f ms init, %.2f ms solve, %.2f ms finish\n", (long)count, launches,
                   launches ? (double)items / launches : 0.0, stat_ms[0] / count, stat_ms[1] / count, stat_ms[2] / count,
                   stat_ms[3] / count);
    }
    return (abip_int)failed.load();
}

}  // extern "C"
