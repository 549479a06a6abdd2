/*
 * abip_gpu.h -- C ABI of the B200-native engine for ABIP's inner ADMM iteration (indirect / pcg=1 path).
 *
 * Three groups of entry points, all extern "C", plain pointers and sizes only:
 *
 *  (1) The reference's linear-system plugin interface, symbol for symbol
 *      (reference src/abip-lp/include/linsys.h:10-91).  Host-pointer semantics are unchanged, so the
 *      unmodified reference solver core links against libabip_gpu.so instead of linsys/indirect.c
 *      (+ linsys/common.c) -- the "drop-in" configuration.
 *
 *  (2) abip_gpu_main/init/solve/finish: same signatures and semantics as the reference's
 *      ABIP(main/init/solve/finish) (src/abip-lp/include/abip.h:119-124, src/abip.c:2056-2422), but all
 *      iterate vectors live in HBM and every inner-loop function runs as a CUDA kernel.  The host reads
 *      one scalar block per ADMM iteration / BB round.
 *
 *  (3) abipgpu_lp_*: the device-resident step functions (2) is built from, exported so that tests can
 *      drive single steps and compare them with the oracle.  New, not in the reference.
 *
 * Types are layout-compatible with the reference's shipped MATLAB build: abip_int = long (-DDLONG,
 * make_abip.m:50-54), abip_float = double (glbopts.h:96-112).
 *
 * Error convention follows the reference: init returns NULL on failure, solve_lin_sys returns < 0 on
 * failure (src/abip.c:1826-1830, 2137-2140); any CUDA error is reported that way and printed to stderr.
 * There is no CPU fallback.
 */
#ifndef ABIP_GPU_H_GUARD
#define ABIP_GPU_H_GUARD

#ifdef __cplusplus
extern "C" {
#endif

typedef long abip_int;
typedef double abip_float;

/* status codes: reference include/glbopts.h:22-31 */
#define ABIP_INFEASIBLE_INACCURATE (-7)
#define ABIP_UNBOUNDED_INACCURATE (-6)
#define ABIP_SIGINT (-5)
#define ABIP_FAILED (-4)
#define ABIP_INDETERMINATE (-3)
#define ABIP_INFEASIBLE (-2)
#define ABIP_UNBOUNDED (-1)
#define ABIP_UNFINISHED (0)
#define ABIP_SOLVED (1)
#define ABIP_SOLVED_INACCURATE (2)

/* CSC matrix: reference linsys/amatrix.h:10-17 */
typedef struct ABIP_A_DATA_MATRIX {
    abip_float *x; /* values, length p[n] */
    abip_int *i;   /* row indices */
    abip_int *p;   /* column pointers, length n+1 */
    abip_int m;
    abip_int n;
} ABIPMatrix;

/* reference include/abip.h:36-79 (field order is the ABI) */
typedef struct ABIP_SETTINGS {
    abip_int normalize;
    abip_int pfeasopt;
    abip_float scale;
    abip_float rho_y;
    abip_float sparsity_ratio;
    abip_int max_ipm_iters;
    abip_int max_admm_iters;
    abip_float max_time;
    abip_float eps;
    abip_float alpha;
    abip_float cg_rate;
    abip_int adaptive;
    abip_float eps_cor;
    abip_float eps_pen;
    abip_float dynamic_sigma;
    abip_float dynamic_x;
    abip_float dynamic_eta;
    abip_int restart_fre;
    abip_int restart_thresh;
    abip_int verbose;
    abip_int warm_start;
    abip_int adaptive_lookback;
    abip_int origin_rescale;
    abip_int pc_ruiz_rescale;
    abip_int qp_rescale;
    abip_int ruiz_iter;
    abip_int hybrid_mu;
    abip_float hybrid_thresh;
    abip_float dynamic_sigma_second;
    abip_int half_update;
    abip_int avg_criterion;
} ABIPSettings;

/* reference include/abip.h:23-34 */
typedef struct ABIP_PROBLEM_DATA {
    abip_int m;
    abip_int n;
    ABIPMatrix *A;
    abip_float *b;
    abip_float *c;
    abip_float sp; /* nnz / (m*n) */
    ABIPSettings *stgs;
} ABIPData;

/* reference include/abip.h:81-86 */
typedef struct ABIP_SOL_VARS {
    abip_float *x;
    abip_float *y;
    abip_float *s;
} ABIPSolution;

/* reference include/abip.h:88-105; times in milliseconds */
typedef struct ABIP_INFO {
    char status[32];
    abip_int status_val;
    abip_int ipm_iter;
    abip_int admm_iter;
    abip_float pobj;
    abip_float dobj;
    abip_float res_pri;
    abip_float res_dual;
    abip_float rel_gap;
    abip_float res_infeas;
    abip_float res_unbdd;
    abip_float setup_time;
    abip_float solve_time;
} ABIPInfo;

/* reference include/abip.h:107-114 */
typedef struct ABIP_SCALING {
    abip_float *D;
    abip_float *E;
    abip_float mean_norm_row_A;
    abip_float mean_norm_col_A;
} ABIPScaling;

typedef struct ABIP_LIN_SYS_WORK ABIPLinSysWork; /* private: holds the device engine (linsys/indirect.h:14-29) */
typedef struct ABIP_GPU_WORK ABIPGpuWork;        /* private: replaces struct ABIP_WORK (abip.h:126-176) */

/* ------------------------------------------------------------------------------------------------
 * (1) linsys plugin -- replaces linsys/indirect.c + linsys/common.c.
 * ---------------------------------------------------------------------------------------------- */
/* indirect.c:282-318: builds CSR(A), CSR(A'), M = 1/diag(AA') in HBM.  NULL on failure. */
ABIPLinSysWork *abip_init_lin_sys_work(const ABIPMatrix *A, const ABIPSettings *stgs);
/* indirect.c:393-434: solves [rho_y I, A; A', -I] sol = b in place (b: host, length >= m+n; s: host warm
 * start, length >= m, may be NULL); iter < 0 requests the 1e-9 tolerance.  0 on success, < 0 on failure. */
abip_int abip_solve_lin_sys(const ABIPMatrix *A, const ABIPSettings *stgs, ABIPLinSysWork *p,
                            abip_float *b, const abip_float *s, abip_int iter);
/* indirect.c:141-203 */
void abip_free_lin_sys_work(ABIPLinSysWork *p);
/* indirect.c:222-242: y += A'x / y += Ax on host vectors (H2D, CSR SpMV kernel, D2H) */
void abip_accum_by_Atrans(const ABIPMatrix *A, ABIPLinSysWork *p, const abip_float *x, abip_float *y);
void abip_accum_by_A(const ABIPMatrix *A, ABIPLinSysWork *p, const abip_float *x, abip_float *y);
/* common.c:44-96 */
abip_int abip_validate_lin_sys(const ABIPMatrix *A);
/* indirect.c:8-33; returned strings are malloc'ed (128 B), caller frees */
char *abip_get_lin_sys_method(const ABIPMatrix *A, const ABIPSettings *stgs);
char *abip_get_lin_sys_summary(ABIPLinSysWork *p, const ABIPInfo *info);
/* common.c:150-594: pc + ruiz (+ origin / qp) equilibration of A in place; allocates scal->D, scal->E */
void abip_normalize_A(ABIPMatrix *A, const ABIPSettings *stgs, ABIPScaling *scal);
void abip_un_normalize_A(ABIPMatrix *A, const ABIPSettings *stgs, const ABIPScaling *scal);
/* common.c:10-41, 100-118 */
void abip_free_A_matrix(ABIPMatrix *A);
abip_int abip_copy_A_matrix(ABIPMatrix **dstp, const ABIPMatrix *src);

/* ------------------------------------------------------------------------------------------------
 * (2) solver entry -- replaces ABIP(main/init/solve/finish), src/abip.c:2056-2422.
 * ---------------------------------------------------------------------------------------------- */
void abip_gpu_set_default_settings(ABIPData *d); /* util.c:288-329 + mex defaults abip_mex.c:320-341 */
ABIPGpuWork *abip_gpu_init(const ABIPData *d, ABIPInfo *info);
abip_int abip_gpu_solve(ABIPGpuWork *w, const ABIPData *d, ABIPSolution *sol, ABIPInfo *info);
void abip_gpu_finish(ABIPGpuWork *w);
abip_int abip_gpu_main(const ABIPData *d, ABIPSolution *sol, ABIPInfo *info);

/* Multi-GPU, one process per GPU (new; the reference is single-process).  Every rank passes the FULL problem to
 * abip_gpu_init_dist; rank r keeps a block of columns of A (a row block of the stored A').  The ranks then exchange
 * the 64-byte CUDA-IPC handles (abip_gpu_comm_export -> all-gather -> abip_gpu_comm_connect) and call abip_gpu_solve
 * collectively; sol->y is complete on every rank, sol->x / sol->s hold this rank's shard [c0, c0+nl) and zeros
 * elsewhere (sum over ranks = full vectors).
 * Not supported by the sharded engine: settings->half_update = 1 (abip_gpu_init_dist prints an error and returns NULL; the
 * half-update pair of src/abip.c:607-700 is implemented in the single-GPU and batch engines only).  Wall-clock decisions
 * (time limit, SIGINT) are taken per rank: give every rank the same limits. */
ABIPGpuWork *abip_gpu_init_dist(const ABIPData *d, ABIPInfo *info, abip_int rank, abip_int world);
abip_int abip_gpu_comm_export(ABIPGpuWork *w, void *handle64);
abip_int abip_gpu_comm_connect(ABIPGpuWork *w, const void *handles /* world x 64 bytes, rank order */);
void abip_gpu_partition(const ABIPGpuWork *w, abip_int *c0, abip_int *nl);
void abip_gpu_column_partition(abip_int n, const abip_int *Ap, abip_int world, abip_int rank, abip_int *c0, abip_int *nl);

/* Batch of independent LPs on one GPU (configs[4]); `concurrency` = problems in flight (one host thread each).
 *   ctas_per_problem <= 1: one CTA per problem; the whole solve (outer loop of ABIP(solve), src/abip.c:2093-2295: inner ADMM
 *     loops, convergence checks, mu rules, re-initialisation, Barzilai-Borwein searches) runs inside one launch of the batch
 *     kernel, the requests of the problems in flight are launched in groups on several streams and every owner is woken
 *     when its CTA has finished (~1.3 x SM count problems in flight is enough).  With verbose != 0, a trace file or past
 *     restart_thresh the host drives the outer loop; time limit and SIGINT are checked between launches;
 *   ctas_per_problem >= 2: each problem on its own stream with a persistent grid of that many CTAs.
 * Returns the number of failed problems (< 0: bad arguments).
 * The reference equivalent is a loop of ABIP(main) calls (one process per core). */
abip_int abip_gpu_batch_main(const ABIPData *const *problems, ABIPSolution *sols, ABIPInfo *infos, abip_int count,
                             abip_int concurrency, abip_int ctas_per_problem);

/* counters of the last abip_gpu_solve on this work (for roofline accounting) */
typedef struct ABIP_GPU_STATS {
    abip_int n_admm_launch;   /* ADMM-iteration kernel launches */
    abip_int n_bb_launch;     /* BB-round kernel launches (2 solves each) */
    abip_int n_solves;        /* solve_lin_sys executions (incl. g = K^-1 h) */
    abip_int n_cg_iters;      /* total CG iterations */
    abip_int n_spmv_A;        /* y = A x passes (each streams CSR(A) once) */
    abip_int n_spmv_AT;       /* y = A' x passes */
    abip_int n_kernel_launches;
    abip_float admm_kernel_ms; /* CUDA-event time inside ADMM-iteration launches */
    abip_float bb_kernel_ms;
    abip_float alg_bytes;      /* algorithmic bytes moved by all launches (formula in DESIGN.md) */
    abip_float h2d_bytes;
    abip_float d2h_bytes;
    abip_float alg_bytes_admm; /* share of alg_bytes inside ADMM-iteration launches */
    abip_float alg_bytes_bb;   /* share inside BB-round launches */
    abip_float solve_event_ms; /* CUDA-event time (engine stream) around the whole abip_gpu_solve */
} ABIPGpuStats;
void abip_gpu_get_stats(const ABIPGpuWork *w, ABIPGpuStats *out);

/* ------------------------------------------------------------------------------------------------
 * (3) device-resident step functions (opaque engine handle).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ABIPGPU_LP abipgpu_lp;

/* vector ids for abipgpu_lp_get_vec / set_vec (lengths: l = m+n+1 unless noted) */
enum {
    ABIPGPU_VEC_U = 0, ABIPGPU_VEC_V = 1, ABIPGPU_VEC_UT = 2, ABIPGPU_VEC_UPREV = 3,
    ABIPGPU_VEC_USUM = 4, ABIPGPU_VEC_VSUM = 5, ABIPGPU_VEC_UAVGC = 6, ABIPGPU_VEC_VAVGC = 7,
    ABIPGPU_VEC_UAVG = 8, ABIPGPU_VEC_VAVG = 9,
    ABIPGPU_VEC_H = 10 /* m+n */, ABIPGPU_VEC_G = 11 /* m+n */, ABIPGPU_VEC_M = 12 /* m */,
    ABIPGPU_VEC_BB_UPREV = 13, ABIPGPU_VEC_BB_VPREV = 14, ABIPGPU_VEC_BB_U = 15, ABIPGPU_VEC_BB_V = 16,
    ABIPGPU_VEC_BB_UNEXT = 17, ABIPGPU_VEC_BB_VNEXT = 18, ABIPGPU_VEC_BB_UT = 19, ABIPGPU_VEC_BB_UTNEXT = 20
};

/* scalar block written by a step (doubles); indices below */
enum {
    ABIPGPU_SC_CG_ITS = 0,    /* CG iterations of the (last) solve in the launch */
    ABIPGPU_SC_CG_ITS2 = 1,   /* second solve of a BB round */
    ABIPGPU_SC_CG_TOL = 2, ABIPGPU_SC_CG_RES = 3,
    /* Q-norm sums for (u,v): src/abip.c:1964-1992 and, D/E-weighted, :385-456 */
    ABIPGPU_SC_S_PR = 4, ABIPGPU_SC_W_AX = 5, ABIPGPU_SC_W_PR = 6, ABIPGPU_SC_BTY = 7, ABIPGPU_SC_UU_Y = 8,
    ABIPGPU_SC_S_DR = 9, ABIPGPU_SC_W_ATYS = 10, ABIPGPU_SC_W_DR = 11, ABIPGPU_SC_CTX = 12,
    ABIPGPU_SC_UU_X = 13, ABIPGPU_SC_VV = 14, ABIPGPU_SC_TAU = 15, ABIPGPU_SC_KAP = 16,
    /* same for (u_avgcon, v_avgcon), valid when ABIPGPU_SC_HAS_AVG != 0 */
    ABIPGPU_SC_AVG_BASE = 20, /* + (index - ABIPGPU_SC_S_PR) */
    ABIPGPU_SC_HAS_AVG = 36,
    /* BB round: adaptive.c:170-178 */
    ABIPGPU_SC_BB_UTUT = 40, ABIPGPU_SC_BB_UTV = 41, ABIPGPU_SC_BB_UU = 42, ABIPGPU_SC_BB_VV = 43,
    ABIPGPU_SC_BB_UV = 44,
    /* mu statistics: abip.c:957-960 */
    ABIPGPU_SC_MIN_XS = 48, ABIPGPU_SC_SUM_XS = 49,
    ABIPGPU_SC_VEC_NORM2 = 50,
    /* device-resident loops of the batch kernel (abipgpu_lp_inner_loop / abipgpu_lp_bb_search) */
    ABIPGPU_SC_LOOP_EXIT = 51, ABIPGPU_SC_LOOP_ITERS = 52, ABIPGPU_SC_LOOP_CG = 53, ABIPGPU_SC_LOOP_AVG = 54,
    ABIPGPU_SC_LOOP_BETA = 56, ABIPGPU_SC_LOOP_ROUNDS = 57,
    /* state of the device-resident outer loop (batch engines) */
    ABIPGPU_SC_LOOP_I = 58, ABIPGPU_SC_LOOP_J = 59, ABIPGPU_SC_LOOP_MU = 60, ABIPGPU_SC_LOOP_SIGMA = 61,
    ABIPGPU_SC_LOOP_GAMMA = 62, ABIPGPU_SC_LOOP_FLAGS = 55 /* final_check | double_check << 1 */,
    ABIPGPU_SC_LOOP_DYN = 45 /* dynamic_sigma */,
    /* batch executor: nanoseconds the item spent on its SM, and its completion flag (host-mapped output only) */
    ABIPGPU_SC_BATCH_NS = 37, ABIPGPU_SC_BATCH_DONE = 38,
    ABIPGPU_SC_COMM_ERR = 63, /* multi-GPU: a peer did not answer within the spin limit */
    ABIPGPU_SC_COUNT = 64
};

/* A: *scaled* CSC (host).  Builds CSR(A), CSR(A'), preconditioner; allocates all work vectors. */
abipgpu_lp *abipgpu_lp_create(abip_int m, abip_int n, const abip_int *Ap, const abip_int *Ai,
                              const abip_float *Ax, const ABIPSettings *stgs, int device);
/* same, but A is UNSCALED: the equilibration of common.c:150-565 runs on the device (bit-identical to
 * abip_normalize_A); D [m], E [n] and the mean row / column norms are returned for normalize_b_c / un_normalize_sol */
abipgpu_lp *abipgpu_lp_create_scaling(abip_int m, abip_int n, const abip_int *Ap, const abip_int *Ai,
                                      const abip_float *Ax, const ABIPSettings *stgs, int device, abip_float *D,
                                      abip_float *E, abip_float *mean_norm_row_A, abip_float *mean_norm_col_A);
void abipgpu_lp_destroy(abipgpu_lp *e);
/* uploads scaled b, c (and D, E, may be NULL when normalize = 0); forms h = [-b; c], solves g = K^-1 h with
 * the iter = -1 tolerance, flips g_x, computes g_th (abip.c:1915-1924).  Returns < 0 on failure. */
int abipgpu_lp_set_problem(abipgpu_lp *e, const abip_float *b, const abip_float *c,
                           const abip_float *D, const abip_float *E);
int abipgpu_lp_cold_start(abipgpu_lp *e, abip_float mu, abip_float beta);              /* abip.c:361-381 */
int abipgpu_lp_outer_prologue(abipgpu_lp *e, int avg_criterion);                        /* abip.c:2117-2129 */
/* one full inner iteration (abip.c:2133-2173): u_prev copy, project_lin_sys, project_barrier + update_dual_vars
 * (or the half_update pair), restart_vars, compute_avg, iterate_Q_norm_resd sums.  j = inner index,
 * k = global ADMM counter (CG tolerance schedule).  sc[ABIPGPU_SC_COUNT] receives the scalar block. */
int abipgpu_lp_admm_iter(abipgpu_lp *e, abip_int j, abip_int k, abip_float mu, abip_float beta,
                         abip_float *sc);
int abipgpu_lp_mu_stats(abipgpu_lp *e, int avg_criterion, abip_float *sc);              /* abip.c:957-960 */
int abipgpu_lp_reinit(abipgpu_lp *e, int indx, abip_float sigma, int avg_criterion);    /* abip.c:996-1075 */
int abipgpu_lp_clamp_v(abipgpu_lp *e);                                                  /* abip.c:2175-2186 */
int abipgpu_lp_bb_begin(abipgpu_lp *e);                                                 /* adaptive.c:86-87 */
/* one lookback round (adaptive.c:89-178).  carry: 0 = first round, 1 = beta changed (:230-242),
 * 2 = keep (u,v) (:243-247). */
int abipgpu_lp_bb_round(abipgpu_lp *e, int carry, abip_int k, abip_float mu, abip_float beta_prev,
                        abip_float *sc);
/* device solve_lin_sys on an engine vector (rhs id is overwritten; warm id < 0 means no warm start) */
int abipgpu_lp_solve_vec(abipgpu_lp *e, int rhs_id, int warm_id, abip_int iter, abip_float *sc);
int abipgpu_lp_get_vec(abipgpu_lp *e, int id, abip_float *host, abip_int len);
int abipgpu_lp_set_vec(abipgpu_lp *e, int id, const abip_float *host, abip_int len);
abip_float abipgpu_lp_g_th(const abipgpu_lp *e);
/* y = A x (trans = 0, x length n, y length m) or y = A' x (trans = 1) on host vectors */
int abipgpu_lp_spmv(abipgpu_lp *e, int trans, const abip_float *x, abip_float *y);
/* launch geometry and SpMV variant chosen from the row-length statistics */
void abipgpu_lp_describe(const abipgpu_lp *e, char *buf, abip_int buflen);
/* The SpMV plan of a CSR matrix for a persistent grid of `ctas` CTAs (host-only, no device needed; for tests and
 * tools).  Outputs: chunk descriptors {row0, nnz0, rows | -(piece slot) - 1, nnz} as 4 ints each (capacity
 * max_chunks; returns the number of chunks, or -needed when the capacity is too small), warp_chunk [ctas * warps + 1],
 * info = {warps per CTA, chunk nonzero limit, chunk row limit, long rows, pieces, lanes per row (log2)}. */
abip_int abipgpu_plan_debug(abip_int nrows, const int *rowptr, abip_int ctas, abip_int deal, int *chunks4,
                            abip_int max_chunks, int *warp_chunk, int *info6);
/* Same, with explicit per-row costs for the cut of the CTA row ranges (row_cost[nrows], NULL: the structural model).
 * This is the path of the measured balance: whole-device engines time the two SpMV passes of the PCG loop per CTA at
 * set-up, rescale the model costs with the measured times and cut the ranges again (lp_engine.cu: tune_balance;
 * ABIP_GPU_TUNE=0 disables it). */
abip_int abipgpu_plan_debug_cost(abip_int nrows, const int *rowptr, abip_int ctas, abip_int deal, const double *row_cost,
                                 int *chunks4, abip_int max_chunks, int *warp_chunk, int *info6);

/* ------------------------------------------------------------------------------------------------
 * (4) ABIP-QCP: min 1/2 x'Qx + c'x  s.t. Ax = b, x in K (SOC, rotated SOC, free, zero, orthant blocks in this
 *     column order).  Entry and structs mirror src/abip-qcp/include/abip.h:67-165 and abip() (source/abip.c:1335);
 *     abip_int of that build is `int` (make_abip_qcp.m does not define DLONG).
 * ---------------------------------------------------------------------------------------------- */
typedef struct ABIP_QCP_MATRIX { /* include/amatrix.h:11-18 (CSC) */
    double *x;
    int *i;
    int *p;
    int m;
    int n;
} ABIPQcpMatrix;

typedef struct ABIP_QCP_CONE { /* include/abip.h:67-76 */
    int *q;
    int qsize;
    int *rq;
    int rqsize;
    int f;
    int z;
    int l;
} ABIPQcpCone;

typedef struct ABIP_QCP_SETTINGS { /* include/abip.h:91-131 (field order is the ABI) */
    int normalize;
    int scale_E;
    int scale_bc;
    double scale;
    double rho_x;
    double rho_y;
    double rho_tau;
    int max_ipm_iters;
    int max_admm_iters;
    double eps;
    double eps_p;
    double eps_d;
    double eps_g;
    double eps_inf;
    double eps_unb;
    double err_dif;
    double alpha;
    double cg_rate;
    int use_indirect;
    int inner_check_period;
    int outer_check_period;
    int verbose;
    int linsys_solver; /* ignored: this engine always uses its m-space Schur PCG (see csrc/qcp_engine.cu) */
    int prob_type;     /* must be 2 or 3 (general QCP) */
    double time_limit; /* seconds */
    double psi;
    int origin_scaling;
    int ruiz_scaling;
    int pc_scaling;
} ABIPQcpSettings;

typedef struct ABIP_QCP_DATA { /* include/abip.h:79-89 */
    int m;
    int n;
    ABIPQcpMatrix *A;
    ABIPQcpMatrix *Q; /* symmetric, both triangles stored; may be NULL */
    double *b;
    double *c;
    double lambda;
    ABIPQcpSettings *stgs;
} ABIPQcpData;

typedef struct ABIP_QCP_INFO { /* include/abip.h:139-157 */
    char status[32];
    int status_val;
    int ipm_iter;
    int admm_iter;
    double pobj;
    double dobj;
    double res_pri;
    double res_dual;
    double rel_gap;
    double res_infeas;
    double res_unbdd;
    double setup_time;
    double solve_time;
    double avg_linsys_time;
    double avg_cg_iters;
} ABIPQcpInfo;

void abip_qcp_gpu_set_default_settings(ABIPQcpData *d); /* source/util.c:203-255 */
/* abip(d, sol, info, K), source/abip.c:1335-1371 */
int abip_qcp_gpu(const ABIPQcpData *d, ABIPSolution *sol, ABIPQcpInfo *info, ABIPQcpCone *K);
/* counters of the last abip_qcp_gpu call in this thread: ADMM-iteration launches, outer CG iterations, inner
 * (H^-1) CG iterations, CUDA-event time inside the launches (ms) */
void abip_qcp_gpu_last_counters(long *n_iter, long *n_cg, long *n_inner, double *kernel_ms);
/* data scaling of the QCP (qcp_config.c:91-491) on host arrays, in place; D [m], E [n] are written */
void abip_qcp_scale_data(ABIPQcpMatrix *A, ABIPQcpMatrix *Q, double *b, double *c, const ABIPQcpCone *K,
                         const ABIPQcpSettings *stgs, double *D, double *E, double *sc_b, double *sc_c);

/* device-resident QCP step functions */
typedef struct ABIPGPU_QCP abipgpu_qcp;
enum {
    ABIPGPU_QSC_CG_ITS = 0, ABIPGPU_QSC_INNER_ITS = 1, ABIPGPU_QSC_TAU_T = 2, ABIPGPU_QSC_CG_RES = 3,
    ABIPGPU_QSC_S_DIFF = 4, ABIPGPU_QSC_S_QU = 5, ABIPGPU_QSC_S_VO = 6, ABIPGPU_QSC_UMU = 7, ABIPGPU_QSC_YB = 8,
    ABIPGPU_QSC_XC = 9, ABIPGPU_QSC_XQX = 10, ABIPGPU_QSC_AXD2 = 11, ABIPGPU_QSC_QXE2 = 12,
    ABIPGPU_QSC_ATYS_E2 = 13, ABIPGPU_QSC_AXB_INF = 14, ABIPGPU_QSC_AXB_D_INF = 15, ABIPGPU_QSC_AX_D_INF = 16,
    ABIPGPU_QSC_RESD_INF = 17, ABIPGPU_QSC_RESD_E_INF = 18, ABIPGPU_QSC_QX_E_INF = 19, ABIPGPU_QSC_TAU = 20,
    ABIPGPU_QSC_VO_TAU = 21, ABIPGPU_QSC_A_COEF = 22,
    ABIPGPU_QSC_COUNT = 32
};
/* A, Q, b, c, D, E: *scaled* data (CSC, int indices); builds CSR copies, preconditioners, the initial point
 * (abip.c:912-992) and r = K^-1[-b; c], a = rho_tau + r'(rho o r) (pre_calculate, abip.c:886-910) */
abipgpu_qcp *abipgpu_qcp_create(int m, int n, const int *Ap, const int *Ai, const double *Ax, const int *Qp,
                                const int *Qi, const double *Qx, const double *b, const double *c, const double *D,
                                const double *E, const int *q, int qsize, const int *rq, int rqsize, int f, int z,
                                int l, double rho_x, double rho_y, double rho_tau, double alpha, double rtol,
                                int device);
void abipgpu_qcp_destroy(abipgpu_qcp *e);
/* one inner iteration (abip.c:1130-1152): projection, barrier subproblem, dual update, conv-check/residual sums */
int abipgpu_qcp_iter(abipgpu_qcp *e, long k, double mu, double beta, double *sc);
/* vec (host, m+n) <- K^-1 vec with K = [rho_y I, A; -A', Q + rho_x I]; warm: y warm start (host, m) or NULL */
int abipgpu_qcp_solve_vec(abipgpu_qcp *e, double *host_vec, const double *host_warm, double rtol, double *sc);
/* the same solve through the reference's own indirect path, restated on the device: n-space operator mat_vec
 * (source/linsys.c:725-750), Jacobi preconditioner init_qcp_precon (qcp_config.c:754-780), qcp_pcg (linsys.c:755-851,
 * |r|_inf stopping rule) inside solve_qcp_linsys (qcp_config.c:826-881).  warm_x: x warm start (host, n) or NULL;
 * tolerance rtol * |reduced rhs|_inf; sc[ABIPGPU_QSC_CG_ITS] = iterations, sc[ABIPGPU_QSC_CG_RES] = final |r|_inf */
int abipgpu_qcp_solve_nspace(abipgpu_qcp *e, double *host_vec, const double *host_warm_x, double rtol, long max_iter,
                             double *sc);
int abipgpu_qcp_get_vec(abipgpu_qcp *e, int id, double *host, long len); /* 0 u, 1 v, 2 u_t, 3 r */
int abipgpu_qcp_set_vec(abipgpu_qcp *e, int id, const double *host, long len);
double abipgpu_qcp_a_coef(const abipgpu_qcp *e);
void abipgpu_qcp_counters(const abipgpu_qcp *e, long *n_iter, long *n_cg, long *n_inner, double *kernel_ms);

#ifdef __cplusplus
}
#endif
#endif
