// lp_engine.cu -- persistent cooperative kernels of the ABIP-LP engine and the device-resident step ABI
// (group (3) of include/abip_gpu.h).  sm_100a only; no cuSPARSE/cuBLAS; no CPU fallback.
#include "lp_engine.h"
#include "lp_logic.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// =========================================================================================================
// Kernels
// =========================================================================================================
struct IterArgs {
    double *u, *v, *ut, *u_prev;
    double *u_sum, *v_sum, *u_avgc, *v_avgc, *u_avg, *v_avg;
    long j, k;
    double mu, beta;
    int half_update;     // stgs->half_update
    int restart_active;  // total_admm_iter >= restart_thresh: maintain u_avg/v_avg (src/abip.c:601-612)
    int restart_fire;    // (j + 1 - fre_old) % fre == 0 and restart_active
    double restart_fre;
};

// Q-norm / residual sums for one (u, v) pair: src/abip.c:1964-1992 (unweighted) and :385-456 (D/E weighted).
// Writes 11 partial sums into reducer slots [slot0, slot0+11).
template <bool DIST>
__device__ __forceinline__ void dev_qnorm_sums(const LpCtx& c, Reducer& R, cg::grid_group& grid, CommState& cs,
                                               const double* u, const double* v, int half, int slot0, bool more) {
    const int m = c.m, n = c.n;
    const double tau = u[m + n];
    const double* y = u;
    const double* x = u + m;
    const double* s = v + m;
    double a[5] = {0, 0, 0, 0, 0};  // S_PR, W_AX, W_PR, BTY, UU_Y
    auto a_epi = [&](int row, double ax) {
        const double bi = __ldg(c.b + row);
        const double pr = fma(-bi, tau, ax);
        double d2 = 1.0;
        if (c.D) { const double d = __ldg(c.D + row); d2 = d * d; }
        const double yi = y[row];
        a[0] = fma(pr, pr, a[0]);
        a[1] = fma(ax * ax, d2, a[1]);
        a[2] = fma(pr * pr, d2, a[2]);
        a[3] = fma(bi, yi, a[3]);
        a[4] = fma(yi, yi, a[4]);
    };
    if constexpr (!DIST) {
        spmv_rows(c.A, x, R.ws, &c.AT, a_epi);
    } else {
        double* slot = comm_vec_slot(c.comm, cs);
        spmv_rows(c.A, x, R.ws, &c.AT, [&](int row, double ax) { slot[row] = ax; });
        comm_sum_vec(c.comm, cs, grid, m, a_epi);
    }
    R.block_store<5>(a, slot0);
    double d[6] = {0, 0, 0, 0, 0, 0};  // S_DR, W_ATYS, W_DR, CTX, UU_X, VV
    spmv_rows(c.AT, y, R.ws, more ? &c.A : nullptr, [&](int row, double aty) {
        const double cj = __ldg(c.c + row);
        const double sj = s[row];
        const double xj = x[row];
        const double dr0 = aty + sj;
        const double dr = fma(-cj, tau, dr0);
        double e2 = 1.0;
        if (c.E) { const double e = __ldg(c.E + row); e2 = e * e; }
        d[0] = fma(dr, dr, d[0]);
        d[1] = fma(dr0 * dr0, e2, d[1]);
        d[2] = fma(dr * dr, e2, d[2]);
        d[3] = fma(cj, xj, d[3]);
        d[4] = fma(xj, xj, d[4]);
        d[5] = fma(sj, sj, d[5]);
    });
    if (half) {  // v_y is nonzero only with half_update
        GRID_STRIDE(i, m) d[5] = fma(v[i], v[i], d[5]);
    }
    R.block_store<6>(d, slot0 + 5);
}

__device__ __forceinline__ void write_qnorm_sc(double* sc, int base, const double* t, const double* u,
                                               const double* v, int lm1) {
    for (int q = 0; q < 11; ++q) sc[base + q] = t[q];
    sc[base + 11] = u[lm1];
    sc[base + 12] = v[lm1];
}

// One full inner ADMM iteration (src/abip.c:2133-2173).
template <bool DIST>
__device__ __forceinline__ void body_admm_iter(const LpCtx& c, const IterArgs& a, unsigned char* smem_raw, bool batched) {
    cg::grid_group grid = cg::this_grid();
    Reducer R = make_reducer(smem_raw, c.partials, batched);
    CommState cs{DIST ? *c.comm.seq : 0ull, false};
    const int m = c.m, lm1 = c.m + c.n, l = lm1 + 1;

#ifdef ABIP_PHASE_TIMING
    if (threadIdx.x == 0 && c.phase_ns) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        c.phase_ns[32 + 2 * VG() * kWarps + VB()] = (double)smid;
    }
#endif
    PHASE_START(tk);
    dev_build_rhs<DIST>(c, R, grid, cs, a.u, a.v, a.ut, a.u_prev);
    PHASE_MARK(c, tk, 0);
    SolveOut so;
    dev_solve_lin_sys<true, DIST>(c, R, grid, cs, a.ut, a.u, a.k, so);
    grid_sync(grid);
    double hd[1];
    hd[0] = dev_finish_epi_dot<DIST>(c, R, grid, cs);
    PHASE_START(tk2);
    const double lam = a.mu / a.beta;
    const double al = c.alpha;
    const double dom = (double)(a.j + 1);

    // project_barrier + update_dual_vars (abip.c:717-748, 567-584) or half_update pair (:663-711),
    // restart_vars (:587-630), compute_avg (:635-659) -- one pass over l
    GRID_STRIDE(i, l) {
        double uti = a.ut[i];
        if (i == lm1) {  // u_t[tau] += u_t[0:l-1].h (abip.c:560); only the owner of this entry may touch it
            uti += hd[0];
            a.ut[lm1] = uti;
        }
        double un, vn;
        if (!a.half_update) {
            if (i < m) {
                vn = a.v[i];
                un = uti - vn;
            } else {
                const double up = a.u_prev[i];
                const double vo = a.v[i];
                const double t = al * uti + (1 - al) * up - vo;
                un = barrier_prox(t, lam);
                vn = vo + (un - al * uti - (1.0 - al) * up);
            }
        } else {
            double vh = a.v[i] + 0.5 * (a.u[i] - uti);
            un = uti - vh;
            if (i >= m) un = barrier_prox(un, lam);
            vn = vh + (un - uti);
        }
        if (a.restart_active) {
            double ua = a.u_avg[i] + un, va = a.v_avg[i] + vn;
            if (a.restart_fire) {
                ua /= a.restart_fre;
                va /= a.restart_fre;
                un = ua;
                vn = va;
                ua = 0.0;
                va = 0.0;
            }
            a.u_avg[i] = ua;
            a.v_avg[i] = va;
        }
        a.u[i] = un;
        const double us = a.u_sum[i] + un;
        a.u_sum[i] = us;
        a.u_avgc[i] = us / dom;
        if (i >= m || a.half_update || a.restart_active) {
            a.v[i] = vn;
            const double vs = a.v_sum[i] + vn;
            a.v_sum[i] = vs;
            a.v_avgc[i] = vs / dom;
        }
    }
    grid_sync(grid);
    PHASE_MARK(c, tk2, 9);

    // iterate_Q_norm_resd sums (abip.c:1951-2051); every 10th inner iteration also on the running average
    const int has_avg = ((a.j + 1) % 10 == 0) ? 1 : 0;
    dev_qnorm_sums<DIST>(c, R, grid, cs, a.u, a.v, a.half_update, 0, has_avg != 0);
    if (has_avg) dev_qnorm_sums<DIST>(c, R, grid, cs, a.u_avgc, a.v_avgc, a.half_update, 11, false);
    release_reducer(R);  // (drain + mbarrier invalidation: k_batch runs several bodies in one launch)
    grid_sync(grid);
    double t[22];
    if (has_avg) R.finish<22>(t);
    else {
        double t11[11];
        R.finish<11>(t11);
        for (int q = 0; q < 11; ++q) t[q] = t11[q];
    }
    if constexpr (DIST) {  // the six x-block sums (slots 5..10, 16..21) are local: sum them over the GPUs
        double x[12];
        for (int q = 0; q < 6; ++q) { x[q] = t[5 + q]; x[6 + q] = has_avg ? t[16 + q] : 0.0; }
        comm_sum_scalars<12>(c.comm, cs, grid, x);
        for (int q = 0; q < 6; ++q) { t[5 + q] = x[q]; if (has_avg) t[16 + q] = x[6 + q]; }
        if (VB() == 0 && threadIdx.x == 0) {
            *c.comm.seq = cs.seq;
            c.sc[ABIPGPU_SC_COMM_ERR] = cs.failed ? 1.0 : 0.0;
        }
    }
    PHASE_MARK(c, tk2, 10);
    if (VB() == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_SC_CG_ITS] = (double)so.its;
        c.sc[ABIPGPU_SC_CG_TOL] = so.tol;
        c.sc[ABIPGPU_SC_CG_RES] = so.res;
        write_qnorm_sc(c.sc, ABIPGPU_SC_S_PR, t, a.u, a.v, lm1);
        c.sc[ABIPGPU_SC_HAS_AVG] = (double)has_avg;
        if (has_avg) write_qnorm_sc(c.sc, ABIPGPU_SC_AVG_BASE, t + 11, a.u_avgc, a.v_avgc, lm1);
    }
}

struct BBArgs {
    double *u_prev, *v_prev, *ut, *u, *v, *ut_next, *u_next, *v_next;
    int carry;
    long k;
    double mu, beta_prev;
};

// One ADMM step as written inside update_adapt_params (src/adaptive.c:89-123 / :126-156): from (up, vp) produce
// (ut, un, vn).  When `dots` is set, also reduces the 5 BB inner products of :158-178 where
// (u_mid, v_mid, v_first) = (up, vp, v_prev of the round).
template <bool DIST>
__device__ __forceinline__ void dev_bb_half(const LpCtx& c, Reducer& R, cg::grid_group& grid, CommState& cs,
                                            const double* up, const double* vp, double* ut, double* un, double* vn,
                                            long k, double lam, const double* v_first, bool dots, int& its) {
    const int m = c.m, lm1 = c.m + c.n, l = lm1 + 1;
    dev_build_rhs<DIST>(c, R, grid, cs, up, vp, ut, nullptr);
    SolveOut so;
    dev_solve_lin_sys<true, DIST>(c, R, grid, cs, ut, up, k, so);
    its = so.its;
    grid_sync(grid);
    double hd[1];
    hd[0] = dev_finish_epi_dot<DIST>(c, R, grid, cs);
    const double al = c.alpha;
    double d[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};  // utut, utv, uu, vv, uv (DIST: +5 = local x part)
    GRID_STRIDE(i, l) {
        double uti = ut[i];
        if (i == lm1) {  // only the owner of the tau entry reads/writes it (no cross-block hazard)
            uti += hd[0];
            ut[lm1] = uti;
        }
        const double upi = up[i], vpi = vp[i];
        double uo, vo;
        if (i < m) {
            uo = uti - vpi;
            vo = vn[i];  // never written by the reference either (calloc'ed zero, adaptive.c:282-285)
        } else {
            const double t = al * uti + (1 - al) * upi - vpi;
            uo = barrier_prox(t, lam);
            vo = vpi + (uo - al * uti - (1 - al) * upi);
            vn[i] = vo;
        }
        un[i] = uo;
        if (dots) {
            // here (up, vp) = (u, v) of the round and (uo, vo) = (u_next, v_next)
            const double dut = 2.0 * vpi + uo - upi - vo - v_first[i];
            const double du = upi - uo;
            const double dv = (uo - upi) * (al - 1.0) + vo - vpi;
            const int o = (DIST && i >= m && i != lm1) ? 5 : 0;  // y and tau entries are replicated
            d[o + 0] = fma(dut, dut, d[o + 0]);
            d[o + 1] = fma(dut, dv, d[o + 1]);
            d[o + 2] = fma(du, du, d[o + 2]);
            d[o + 3] = fma(dv, dv, d[o + 3]);
            d[o + 4] = fma(du, dv, d[o + 4]);
        }
    }
    if (dots) {
        if constexpr (DIST) R.block_store<10>(d);
        else {
            double d5[5] = {d[0], d[1], d[2], d[3], d[4]};
            R.block_store<5>(d5);
        }
    }
    grid_sync(grid);
    if (dots) {
        if constexpr (DIST) {
            R.finish<10>(d);
            double x[5] = {d[5], d[6], d[7], d[8], d[9]};
            comm_sum_scalars<5>(c.comm, cs, grid, x);
            for (int q = 0; q < 5; ++q) d[q] += x[q];
        } else {
            double d5[5];
            R.finish<5>(d5);
            for (int q = 0; q < 5; ++q) d[q] = d5[q];
        }
        if (VB() == 0 && threadIdx.x == 0)
            for (int q = 0; q < 5; ++q) c.sc[ABIPGPU_SC_BB_UTUT + q] = d[q];
    }
}

// One lookback round of the Barzilai-Borwein search (src/adaptive.c:89-178).
template <bool DIST>
__device__ __forceinline__ void body_bb_round(const LpCtx& c, const BBArgs& a, unsigned char* smem_raw, bool batched) {
    cg::grid_group grid = cg::this_grid();
    Reducer R = make_reducer(smem_raw, c.partials, batched);
    CommState cs{DIST ? *c.comm.seq : 0ull, false};
    const int m = c.m, l = c.m + c.n + 1;
    const double lam = a.mu / a.beta_prev;
    if (a.carry) {  // state hand-over of the previous round (adaptive.c:230-247)
        GRID_STRIDE(i, l) {
            const double ui = a.u[i];
            a.u_prev[i] = ui;
            a.v_prev[i] = (a.carry == 1 && i >= m) ? lam / ui : a.v[i];
        }
        grid_sync(grid);
    }
    int its1, its2;
    dev_bb_half<DIST>(c, R, grid, cs, a.u_prev, a.v_prev, a.ut, a.u, a.v, a.k, lam, nullptr, false, its1);
    dev_bb_half<DIST>(c, R, grid, cs, a.u, a.v, a.ut_next, a.u_next, a.v_next, a.k, lam, a.v_prev, true, its2);
    release_reducer(R);  // no asynchronous copy may be outstanding when the CTA exits
    if (DIST && VB() == 0 && threadIdx.x == 0) {
        *c.comm.seq = cs.seq;
        c.sc[ABIPGPU_SC_COMM_ERR] = cs.failed ? 1.0 : 0.0;
    }
    if (VB() == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_SC_CG_ITS] = (double)its1;
        c.sc[ABIPGPU_SC_CG_ITS2] = (double)its2;
    }
}

// solve_lin_sys on a device vector; post_g: g_x *= -1 and g_th = h.g (src/abip.c:1922-1924)
template <bool DIST>
__device__ __forceinline__ void body_solve_vec(const LpCtx& c, double* b, const double* s, long iter, int post_g,
                                               unsigned char* smem_raw, bool batched) {
    cg::grid_group grid = cg::this_grid();
    Reducer R = make_reducer(smem_raw, c.partials, batched);
    CommState cs{DIST ? *c.comm.seq : 0ull, false};
    SolveOut so;
    spmv_prefetch(c.A, R.ws);
    dev_solve_lin_sys<false, DIST>(c, R, grid, cs, b, s, iter, so);
    R.ws.drain();
    if (post_g) {
        grid_sync(grid);
        const int m = c.m, lm1 = c.m + c.n;
        double a[2] = {0.0, 0.0};
        GRID_STRIDE(i, lm1) {
            double gi = b[i];
            if (i >= m) { gi = -gi; b[i] = gi; }
            const int slot = (DIST && i >= m) ? 1 : 0;
            a[slot] = fma(__ldg(c.h + i), gi, a[slot]);
        }
        R.block_store<2>(a);
        grid_sync(grid);
        R.finish<2>(a);
        if constexpr (DIST) {
            double x[1] = {a[1]};
            comm_sum_scalars<1>(c.comm, cs, grid, x);
            a[1] = x[0];
        }
        if (VB() == 0 && threadIdx.x == 0) c.sc[ABIPGPU_SC_VEC_NORM2] = a[0] + a[1];
    }
    if (DIST && VB() == 0 && threadIdx.x == 0) {
        *c.comm.seq = cs.seq;
        c.sc[ABIPGPU_SC_COMM_ERR] = cs.failed ? 1.0 : 0.0;
    }
    if (VB() == 0 && threadIdx.x == 0) {
        c.sc[ABIPGPU_SC_CG_ITS] = (double)so.its;
        c.sc[ABIPGPU_SC_CG_TOL] = so.tol;
        c.sc[ABIPGPU_SC_CG_RES] = so.res;
    }
    release_reducer(R);  // (a kernel that goes on with other bodies initialises the warp's mbarrier again)
}
template <bool DIST>
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM)
    k_solve_vec(LpCtx c, double* b, const double* s, long iter, int post_g) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    body_solve_vec<DIST>(c, b, s, iter, post_g, smem_raw, false);
}

// y (+)= A x as a stand-alone launch (plugin accum_by_A / accum_by_Atrans and tests)
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM)
    k_spmv(Csr A, const double* x, double* y, int accumulate) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, nullptr);
    spmv_rows(A, x, R.ws, nullptr, [&](int row, double a) { y[row] = accumulate ? y[row] + a : a; });
    R.ws.drain();
}

// Measured balance (tune_balance() below): `reps` iterations of the two SpMV passes of the PCG loop (A' then A with the
// seven-sum epilogue, then the vector update, same barriers) on the engine's own matrices; every CTA reports the SM cycles
// between the start of a pass and the moment its last warp has finished its rows.  The first iteration is a warm-up.
// The PCG workspace it runs on is all zeros (alpha = 0 keeps it so); only the timing matters.
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_tune(LpCtx c, double* t_out /* [3 + 8][G] */, int reps) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Reducer R = make_reducer(smem_raw, c.partials);
    const int m = c.m;
    long long acc_at = 0, acc_a = 0;
    // sections of one iteration as seen by thread 0 of every CTA (SM cycles): [0] A' pass, [1] wait at barrier 1, [2] A pass,
    // [3] block_store, [4] wait at barrier 2, [5] finish, [6] vector update, [7] wait at barrier 3
    long long sec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double sink = 0.0;
    spmv_prefetch(c.AT, R.ws);
    grid_sync(grid);
    for (int rep = 0; rep < reps; ++rep) {
        const long long t0 = clock64();
        spmv_rows(c.AT, c.p, R.ws, &c.A, [&](int row, double a) { c.tmp[row] = a; });
        __syncthreads();
        const long long t1 = clock64();
        grid_sync(grid);
        double d[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        const long long t2 = clock64();
        spmv_rows(c.A, c.tmp, R.ws, &c.AT, [&](int row, double a) {
            const double pi = c.p[row], ri = c.r[row], Mi = __ldg(c.M + row);
            const double gp = fma(c.rho_y, pi, a);
            c.Gp[row] = gp;
            const double zi = Mi * ri, mg = Mi * gp;
            d[0] = fma(pi, gp, d[0]);
            d[1] = fma(zi, gp, d[1]);
            d[2] = fma(mg, gp, d[2]);
            d[3] = fma(ri, gp, d[3]);
            d[4] = fma(gp, gp, d[4]);
            d[5] = fma(zi, ri, d[5]);
            d[6] = fma(ri, ri, d[6]);
        });
        __syncthreads();
        const long long t3 = clock64();
        R.block_store<7>(d);
        const long long t4 = clock64();
        grid_sync(grid);
        const long long t5 = clock64();
        R.finish<7>(d);
        const long long t6 = clock64();
        sink += d[0] + d[1] + d[2] + d[3] + d[4] + d[5] + d[6];
        const double alpha = 0.0 * sink;
        GRID_STRIDE(i, m) {
            const double pi = c.p[i];
            const double ri = fma(-alpha, c.Gp[i], c.r[i]);
            c.r[i] = ri;
            c.p[i] = fma(1.0, pi, alpha * (__ldg(c.M + i) * ri));
        }
        __syncthreads();
        const long long t7 = clock64();
        grid_sync(grid);
        const long long t8 = clock64();
        if (rep > 0) {
            acc_at += t1 - t0;
            acc_a += t3 - t2;
            sec[0] += t1 - t0; sec[1] += t2 - t1; sec[2] += t3 - t2; sec[3] += t4 - t3;
            sec[4] += t5 - t4; sec[5] += t6 - t5; sec[6] += t7 - t6; sec[7] += t8 - t7;
        }
    }
    release_reducer(R);
    if (threadIdx.x == 0) {
        t_out[VB()] = (double)acc_at;
        t_out[VG() + VB()] = (double)acc_a;
        t_out[2 * VG() + VB()] = sink;
#pragma unroll
        for (int k = 0; k < 8; ++k) t_out[(3 + k) * VG() + VB()] = (double)sec[k];
    }
}

// min / sum of u_i v_i over the (x, tau) tail: update_barrier_dynamic, src/abip.c:957-960
struct MuArgs {
    const double *u, *v;
};
__device__ __forceinline__ void body_mu_stats(const double* u, const double* v, int m, int l, double* partials, double* sc,
                                              const Comm& comm, bool batched) {
    cg::grid_group grid = cg::this_grid();
    __shared__ double s_sum[kWarps], s_min[kWarps];
    vgrid_init(batched);
    double sum = 0.0, mn = 1e10;
    GRID_STRIDE(i, l) {
        if (i >= m) {
            const double xs = u[i] * v[i];
            sum += xs;
            mn = fmin(mn, xs);
        }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, off);
        mn = fmin(mn, __shfl_down_sync(0xffffffffu, mn, off));
    }
    if (lane == 0) { s_sum[w] = sum; s_min[w] = mn; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0, q = 1e10;
        for (int i = 0; i < kWarps; ++i) { s += s_sum[i]; q = fmin(q, s_min[i]); }
        partials[VB()] = s;
        partials[VG() + VB()] = q;
    }
    grid_sync(grid);
    double s = 0.0, q = 1e10;
    if (comm.G > 1 || (VB() == 0 && threadIdx.x == 0))
        for (int i = 0; i < VG(); ++i) { s += __ldcg(partials + i); q = fmin(q, __ldcg(partials + VG() + i)); }
    if (comm.G > 1) {
        // the tau entry (index l-1) is replicated: take it out of the local sums, exchange, add it back once
        const double xt = u[l - 1] * v[l - 1];
        CommState cs{*comm.seq, false};
        double* mine = comm.scal[comm.rank] + ((cs.seq + 1) & 1ull) * kCommScalars;
        if (VB() == 0 && threadIdx.x == 0) {
            // local min over the x shard only (recompute without tau is not possible from the partials: min is
            // idempotent, so including tau on every rank is harmless); the sum must drop it
            mine[0] = s - xt;
            mine[1] = q;
        }
        grid_sync(grid);
        comm_exchange(comm, cs, grid);
        const long off = (long)(cs.seq & 1ull) * kCommScalars;
        s = xt;
        q = 1e10;
        for (int r = 0; r < comm.G; ++r) {
            s += __ldcv(comm.scal[r] + off);
            q = fmin(q, __ldcv(comm.scal[r] + off + 1));
        }
        if (VB() == 0 && threadIdx.x == 0) {
            *comm.seq = cs.seq;
            sc[ABIPGPU_SC_COMM_ERR] = cs.failed ? 1.0 : 0.0;
        }
    }
    if (VB() == 0 && threadIdx.x == 0) {
        sc[ABIPGPU_SC_SUM_XS] = s;
        sc[ABIPGPU_SC_MIN_XS] = q;
    }
}

// ---- launch wrappers -------------------------------------------------------------------------------------------
template <bool DIST>
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_admm_iter(LpCtx c, IterArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    body_admm_iter<DIST>(c, a, smem_raw, false);
}
template <bool DIST>
__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_bb_round(LpCtx c, BBArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    body_bb_round<DIST>(c, a, smem_raw, false);
}
__global__ void __launch_bounds__(kBlock) k_mu_stats(const double* u, const double* v, int m, int l, double* partials,
                                                     double* sc, Comm comm) {
    body_mu_stats(u, v, m, l, partials, sc, comm, false);
}

// Batched step (BASELINE.json configs[4]: thousands of small independent LPs): ONE launch advances every problem of the
// batch by one step of its own kind -- CTA b runs item b (an ADMM iteration, a BB round or the mu statistics of one
// engine) as block 0 of a one-block virtual grid, then copies the engine's scalar block to the host-mapped output.
// Replaces one launch + one copy + one synchronisation per problem and step by one launch + one synchronisation per
// batch and step (the per-context driver lock serialised the per-problem calls at ~13 us each).
enum { BATCH_ADMM = 0, BATCH_BB = 1, BATCH_MU = 2, BATCH_INNER = 3, BATCH_BBSEARCH = 4, BATCH_SOLVE = 5 };
struct BBSearchArgs {
    int lookback;
    double eps_cor, eps_pen;
};
// Deferred vector operations of a batch engine (cold start, outer-iteration prologue, re-initialisation, clamp, BB
// hand-over, restart copy): the host functions only record them, the next batched step of the problem executes them
// in order before its own work -- no launch, no copy, no synchronisation of their own (they were ~8 driver calls per
// outer iteration and problem).
enum { PRE_COLD = 1, PRE_PROLOGUE, PRE_REINIT, PRE_CLAMP, PRE_BB_BEGIN, PRE_RESTART_COPY };
struct PreOp {
    int code, i0, i1, pad_;
    double d0, d1;
};
constexpr int kMaxPre = 8;
struct BatchItem {
    LpCtx c;
    int kind;
    int n_pre;
    IterArgs it;
    BBArgs bb;
    MuArgs mu;
    LpSolveArgs solve;  // BATCH_INNER reads solve.in only
    BBSearchArgs search;
    int resident_bytes;  // shared memory needed to keep A and A' resident (0: not eligible)
    int solve_g;         // 1: g = K^-1 h and g_th = h.g (abipgpu_lp_set_problem) are still to be computed, before the step itself
    PreOp pre[kMaxPre];
    double* vec[21];
};
__device__ __noinline__ void dev_pre_op(const BatchItem& it, const PreOp op) {
    const int m = it.c.m, l = it.c.m + it.c.n + 1;
    {
        if (op.code == PRE_COLD) {  // cold_start_vars, src/abip.c:361-381
            const double val = sqrt(op.d0 / op.d1);
            double *u = it.vec[ABIPGPU_VEC_U], *v = it.vec[ABIPGPU_VEC_V];
            for (int i = threadIdx.x; i < l; i += kBlock) {
                u[i] = (i < m) ? 0.0 : val;
                v[i] = (i < m) ? 0.0 : val;
            }
        } else if (op.code == PRE_PROLOGUE) {  // start of an outer iteration (abipgpu_lp_outer_prologue)
            double *us = it.vec[ABIPGPU_VEC_USUM], *vs = it.vec[ABIPGPU_VEC_VSUM], *ua = it.vec[ABIPGPU_VEC_UAVG],
                   *va = it.vec[ABIPGPU_VEC_VAVG];
            for (int i = threadIdx.x; i < l; i += kBlock) {
                us[i] = 0.0;
                vs[i] = 0.0;
                ua[i] = 0.0;
                va[i] = 0.0;
            }
            if (op.i0) {
                double *u = it.vec[ABIPGPU_VEC_U], *v = it.vec[ABIPGPU_VEC_V];
                const double *uc = it.vec[ABIPGPU_VEC_UAVGC], *vc = it.vec[ABIPGPU_VEC_VAVGC];
                for (int i = threadIdx.x; i < l; i += kBlock) {
                    u[i] = uc[i];
                    v[i] = vc[i];
                }
            }
        } else if (op.code == PRE_REINIT) {  // reinitialize_vars, src/abip.c:996-1075
            double* u = it.vec[op.i1 ? ABIPGPU_VEC_UAVGC : ABIPGPU_VEC_U];
            double* v = it.vec[op.i1 ? ABIPGPU_VEC_VAVGC : ABIPGPU_VEC_V];
            const int indx = op.i0;
            const double sigma = op.d0;
            for (int i = m + threadIdx.x; i < l; i += kBlock) {
                if (indx == 0) {
                    if (u[i] > v[i]) v[i] = sigma * v[i];
                    else u[i] = sigma * u[i];
                } else {
                    const double f = (indx == 1) ? sqrt(sigma) : sqrt(1.0 / sigma);
                    u[i] = f * u[i];
                    v[i] = f * v[i];
                }
            }
        } else if (op.code == PRE_CLAMP) {  // src/abip.c:2175-2186
            double* v = it.vec[ABIPGPU_VEC_V];
            for (int i = threadIdx.x; i < l; i += kBlock)
                if (v[i] < 0) v[i] = 1e-6;
        } else if (op.code == PRE_BB_BEGIN) {
            double *up = it.vec[ABIPGPU_VEC_BB_UPREV], *vp = it.vec[ABIPGPU_VEC_BB_VPREV];
            const double *u = it.vec[ABIPGPU_VEC_U], *v = it.vec[ABIPGPU_VEC_V];
            for (int i = threadIdx.x; i < l; i += kBlock) {
                up[i] = u[i];
                vp[i] = v[i];
            }
        } else if (op.code == PRE_RESTART_COPY) {
            double *ua = it.vec[ABIPGPU_VEC_UAVG], *va = it.vec[ABIPGPU_VEC_VAVG];
            const double *us = it.vec[ABIPGPU_VEC_USUM], *vs = it.vec[ABIPGPU_VEC_VSUM];
            for (int i = threadIdx.x; i < l; i += kBlock) {
                ua[i] = us[i];
                va[i] = vs[i];
            }
        }
        __syncthreads();
    }
}
__device__ __forceinline__ void apply_pre_ops(const BatchItem& it) {
    for (int q = 0; q < it.n_pre; ++q) dev_pre_op(it, it.pre[q]);
}
// Shared-memory-resident matrices of a batch problem: [A: val f64 | A': val f64 | A: ptr i32 | A': ptr i32 | A: idx u16 |
// A': idx u16] behind the reducer / barrier / plan-cache area (the TMA stages are not used in this mode).
__host__ __device__ inline size_t resident_bytes_for(long m, long n, long nnz) {
    return kStageOff + 16 * (size_t)nnz + 4 * (size_t)(m + n + 2) + 4 * (size_t)nnz + 64;
}
__device__ __forceinline__ void load_resident(Csr& A, unsigned char*& p) {
    const int nnz = __ldg(A.ptr + A.nrows);
    double* v = reinterpret_cast<double*>(p);
    for (int i = threadIdx.x; i < nnz; i += kBlock) v[i] = __ldg(A.val + i);
    A.sm_val = v;
    p += 8 * (size_t)nnz;
}
__device__ __forceinline__ void load_resident_ptr(Csr& A, unsigned char*& p) {
    int* q = reinterpret_cast<int*>(p);
    for (int i = threadIdx.x; i <= A.nrows; i += kBlock) q[i] = __ldg(A.ptr + i);
    A.sm_ptr = q;
    p += 4 * (size_t)(A.nrows + 1);
}
__device__ __forceinline__ void load_resident_idx(Csr& A, unsigned char*& p) {
    const int nnz = __ldg(A.ptr + A.nrows);
    unsigned short* q = reinterpret_cast<unsigned short*>(p);
    for (int i = threadIdx.x; i < nnz; i += kBlock) q[i] = (unsigned short)__ldg(A.idx + i);
    A.sm_idx = q;
    p += 2 * (size_t)((nnz + 7) & ~7);
}

__global__ void __launch_bounds__(kBlock, ABIP_MIN_BLOCKS_PER_SM) k_batch(const BatchItem* items, double* sc_out, int resident,
                                                                          double seq) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ __align__(16) BatchItem s_item;
    unsigned long long t_begin = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    {
        const int* src = reinterpret_cast<const int*>(items + blockIdx.x);
        int* dst = reinterpret_cast<int*>(&s_item);
        for (int i = threadIdx.x; i < (int)(sizeof(BatchItem) / sizeof(int)); i += kBlock) dst[i] = src[i];
    }
    __syncthreads();
    if (resident) {
        // this launch has the shared memory for it: the CTA copies both matrices of its problem once and every SpMV pass of
        // every iteration of the step reads them from there (re-streaming 240 KB per CG iteration for 296 problems at once
        // is what bound the batch: 7 TB/s of HBM traffic)
        Csr A = s_item.c.A, AT = s_item.c.AT;
        unsigned char* p = smem_raw + kStageOff;
        load_resident(A, p);
        load_resident(AT, p);
        load_resident_ptr(A, p);
        load_resident_ptr(AT, p);
        p = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
        load_resident_idx(A, p);
        load_resident_idx(AT, p);
        __syncthreads();
        if (threadIdx.x == 0) {
            s_item.c.A = A;
            s_item.c.AT = AT;
        }
        __syncthreads();
    }
    const BatchItem& it = s_item;
    apply_pre_ops(it);
    if (it.solve_g) {
        // g = K^-1 h, g_x *= -1, g_th = h.g (src/abip.c:1915-1924) inside the first step of the problem, on the resident
        // matrices: as a launch of its own (57 KB of shared memory) it could not share an SM with a resident k_batch CTA and
        // waited for a free one
        body_solve_vec<false>(it.c, it.vec[ABIPGPU_VEC_G], nullptr, -1, 1, smem_raw, true);
        __syncthreads();
        if (threadIdx.x == 0) s_item.c.g_th = it.c.sc[ABIPGPU_SC_VEC_NORM2];
        __syncthreads();
    }
    if (it.kind == BATCH_ADMM) body_admm_iter<false>(it.c, it.it, smem_raw, true);
    else if (it.kind == BATCH_BB) body_bb_round<false>(it.c, it.bb, smem_raw, true);
    else if (it.kind == BATCH_INNER || it.kind == BATCH_BBSEARCH || it.kind == BATCH_SOLVE) {
        // Device-resident loops (decisions of the host loop taken redundantly by every thread of the CTA from the scalar block,
        // lp_logic.h):  BATCH_INNER = the inner ADMM loop of one outer iteration (src/abip.c:2131-2214), BATCH_BBSEARCH = one
        // Barzilai-Borwein search (src/adaptive.c:34-256), BATCH_SOLVE = the OUTER loop (src/abip.c:2093-2295): inner loop,
        // convergence check, mu rule, re-initialisation, BB search, next outer iteration ... until the solve ends, the
        // launch cap is reached or the restart bookkeeping needs the host.  One copy of the ADMM step and of the BB round
        // serves all three kinds.
        const LpSolveArgs& S = it.solve;
        const LpInnerArgs& L = S.in;
        const bool only_bb = it.kind == BATCH_BBSEARCH, whole = it.kind == BATCH_SOLVE;
        IterArgs a = it.it;
        BBArgs b = it.bb;
        LpMuState ms{L.mu, S.sigma, L.gamma, S.dynamic_sigma, L.final_check, S.double_check};
        double beta = L.beta, cg = 0.0, bb_beta = 0.0;
        long i = L.ipm_iter, j = L.j0, k = L.k0, done = 0;
        int avg = L.avg_in, code = LP_INNER_CONTINUE, rounds = 0;
        bool resume = !whole || S.resume_inner != 0;
        const double spmin = fmin(S.mp.sp, S.mp.sparsity_ratio);
        for (;; ++i) {
            if (!only_bb) {
                if (whole && i >= L.max_ipm_iters) { code = LP_SOLVE_IPM; break; }
                const long j_end = whole ? lp_inner_stopper(spmin, ms.mu, L.max_admm_iters) : L.j_end;
                if (!resume) {  // start of an outer iteration (abipgpu_lp_outer_prologue)
                    dev_pre_op(it, PreOp{PRE_PROLOGUE, avg, 0, 0, 0.0, 0.0});
                    j = 0;
                }
                resume = false;
                int icode = LP_INNER_STOPPER;
                while (j < j_end) {
                    if (k >= L.restart_thresh) { icode = LP_INNER_HOST; break; }
                    if (done >= L.cap) { icode = LP_INNER_CONTINUE; break; }
                    a.j = j;
                    a.k = k;
                    a.mu = ms.mu;
                    a.beta = beta;
                    a.restart_active = 0;
                    a.restart_fire = 0;
                    body_admm_iter<false>(it.c, a, smem_raw, true);
                    __syncthreads();
                    const double* sc = it.c.sc;
                    k += 1;
                    done += 1;
                    cg += sc[ABIPGPU_SC_CG_ITS];
                    const double q = lp_qnorm_decide(sc, (double)L.max_admm_iters, &avg);
                    if (q < ms.gamma * ms.mu) {
                        if (L.half_update) {  // src/abip.c:2175-2186
                            double* v = a.v;
                            for (int t = threadIdx.x; t < it.c.m + it.c.n + 1; t += kBlock)
                                if (v[t] < 0) v[t] = 1e-6;
                        }
                        icode = LP_INNER_CONVERGED;
                        break;
                    }
                    if (ms.final_check) {
                        LpResid r;
                        lp_calc_residuals(L.rin, sc, avg, &r);
                        const int status = lp_has_converged(L.eps, L.pfeasopt, &r, i, k);
                        if (status != 0 || k + 1 >= L.max_admm_iters || i + 1 >= L.max_ipm_iters) { icode = LP_INNER_FINISHED; break; }
                    }
                    ++j;
                    icode = LP_INNER_STOPPER;
                    __syncthreads();  // every thread has read the scalar block before the next iteration rewrites it
                }
                __syncthreads();
                if (!whole || icode == LP_INNER_HOST || icode == LP_INNER_CONTINUE || icode == LP_INNER_FINISHED) {
                    code = icode;
                    break;
                }
                // after the inner loop (src/abip.c:2216-2277)
                if (ms.mu < L.eps) ms.final_check = 1;
                LpResid r;
                lp_calc_residuals(L.rin, it.c.sc, avg, &r);
                const int status = lp_has_converged(L.eps, L.pfeasopt, &r, i, k);
                if (status != 0 || k + 1 >= L.max_admm_iters) { code = LP_SOLVE_DONE; break; }
                const int rule = lp_mu_rule(&ms, S.mp);
                if (rule == 1) lp_update_barrier(&ms, S.mp, r);
                else if (rule == 2) lp_update_barrier_dynamic_2(&ms, S.mp);
                else if (rule == 3) {
                    __syncthreads();
                    body_mu_stats(it.vec[avg ? ABIPGPU_VEC_UAVGC : ABIPGPU_VEC_U], it.vec[avg ? ABIPGPU_VEC_VAVGC : ABIPGPU_VEC_V],
                                  it.c.m, it.c.m + it.c.n + 1, it.c.partials, it.c.sc, it.c.comm, true);
                    __syncthreads();
                    if (lp_update_barrier_dynamic(&ms, S.mp, it.c.sc[ABIPGPU_SC_MIN_XS], it.c.sc[ABIPGPU_SC_SUM_XS]) < 0) {
                        code = LP_SOLVE_FAIL;
                        break;
                    }
                }
                __syncthreads();
                dev_pre_op(it, PreOp{PRE_REINIT, 0, avg, 0, ms.sigma, 0.0});
                if (!S.adaptive) continue;
                if (S.adaptive_lookback <= 0) { code = LP_SOLVE_FAIL; break; }
                dev_pre_op(it, PreOp{PRE_REINIT, 1, avg, 0, ms.sigma, 0.0});
                beta = 1.0;
                dev_pre_op(it, PreOp{PRE_BB_BEGIN, 0, 0, 0, 0.0, 0.0});
            }
            {  // the whole Barzilai-Borwein search: lookback rounds + the safeguarded step
                double beta_prev = 1.0;
                int carry = 0;
                bb_beta = 0.0;
                if (!only_bb) {
                    b.k = k;
                    b.mu = ms.mu;
                }
                const int lookback = only_bb ? it.search.lookback : S.adaptive_lookback;
                const double eps_cor = only_bb ? it.search.eps_cor : S.eps_cor, eps_pen = only_bb ? it.search.eps_pen : S.eps_pen;
                for (int t = 0; t < lookback; ++t) {
                    b.carry = carry;
                    b.beta_prev = beta_prev;
                    body_bb_round<false>(it.c, b, smem_raw, true);
                    __syncthreads();
                    const double* sc = it.c.sc;
                    cg += sc[ABIPGPU_SC_CG_ITS] + sc[ABIPGPU_SC_CG_ITS2];
                    ++rounds;
                    const int action = lp_bb_step(sc, eps_cor, eps_pen, &beta_prev, &bb_beta);
                    if (action == 0) break;
                    carry = action;
                    __syncthreads();
                }
                __syncthreads();
            }
            if (only_bb) break;
            beta = bb_beta;
            dev_pre_op(it, PreOp{PRE_REINIT, 2, avg, 0, ms.sigma, 0.0});
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double* sc = it.c.sc;
            sc[ABIPGPU_SC_LOOP_EXIT] = (double)code;
            sc[ABIPGPU_SC_LOOP_ITERS] = (double)done;
            sc[ABIPGPU_SC_LOOP_CG] = cg;
            sc[ABIPGPU_SC_LOOP_AVG] = (double)avg;
            sc[ABIPGPU_SC_LOOP_BETA] = only_bb ? bb_beta : beta;
            sc[ABIPGPU_SC_LOOP_ROUNDS] = (double)rounds;
            sc[ABIPGPU_SC_LOOP_I] = (double)i;
            sc[ABIPGPU_SC_LOOP_J] = (double)j;
            sc[ABIPGPU_SC_LOOP_MU] = ms.mu;
            sc[ABIPGPU_SC_LOOP_SIGMA] = ms.sigma;
            sc[ABIPGPU_SC_LOOP_GAMMA] = ms.gamma;
            sc[ABIPGPU_SC_LOOP_FLAGS] = (double)(ms.final_check | (ms.double_check << 1));
            sc[ABIPGPU_SC_LOOP_DYN] = ms.dynamic_sigma;
        }
    } else body_mu_stats(it.mu.u, it.mu.v, it.c.m, it.c.m + it.c.n + 1, it.c.partials, it.c.sc, it.c.comm, true);
    __syncthreads();
    // scalar block of the item to the host-mapped output; then its time on the SM and, last, the completion flag (the launch
    // sequence number): the executor releases the owner of an item as soon as the flag shows, not when the whole launch ends
    double* out = sc_out + (size_t)blockIdx.x * ABIPGPU_SC_COUNT;
    if (threadIdx.x < ABIPGPU_SC_COUNT && threadIdx.x != ABIPGPU_SC_BATCH_DONE && threadIdx.x != ABIPGPU_SC_BATCH_NS)
        out[threadIdx.x] = it.c.sc[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        out[ABIPGPU_SC_BATCH_NS] = (double)(t_end - t_begin);
        __threadfence_system();
        *(volatile double*)(out + ABIPGPU_SC_BATCH_DONE) = seq;
    }
}

// reinitialize_vars, src/abip.c:996-1075
__global__ void k_reinit(double* u, double* v, int m, int l, int indx, double sigma) {
    const int i = m + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= l) return;
    if (indx == 0) {
        if (u[i] > v[i]) v[i] = sigma * v[i];
        else u[i] = sigma * u[i];
    } else {
        const double f = (indx == 1) ? sqrt(sigma) : sqrt(1.0 / sigma);
        u[i] = f * u[i];
        v[i] = f * v[i];
    }
}

// half_update inner-loop exit (src/abip.c:2175-2186): negative entries of v (rounding-level) are reset to 1e-6
__global__ void k_clamp_v(double* v, int l) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < l && v[i] < 0) v[i] = 1e-6;
}

// cold_start_vars, src/abip.c:361-381
__global__ void k_cold_start(double* u, double* v, int m, int l, double val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= l) return;
    u[i] = (i < m) ? 0.0 : val;
    v[i] = (i < m) ? 0.0 : val;
}

// h = [-b; c], g = h (src/abip.c:1915-1919)
__global__ void k_build_h(const double* b, const double* c, double* h, double* g, int m, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m + n) return;
    const double v = (i < m) ? -b[i] : c[i - m];
    h[i] = v;
    g[i] = v;
}

// M = 1 / diag(A A') from CSR(A) (linsys/indirect.c:36-79)
__global__ void k_precond(Csr A, double* M) {
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= A.nrows) return;
    double s = 0.0;
    for (int k = A.ptr[row]; k < A.ptr[row + 1]; ++k) s = fma(A.val[k], A.val[k], s);
    M[row] = 1.0 / s;
}


// =========================================================================================================
// Equilibration of A on the device (reference linsys/common.c:150-565).  Works on the CSC
// values (= CSR(A') values); CSR(A) values are gathered afterwards through the transposition permutation.  Sums run
// sequentially in the reference's order with separate multiply and add (no FMA contraction), so D, E and the scaled
// matrix are bit-identical to the host implementation abip_normalize_A.
// kind: 0 pc (sqrt of 1-norm), 1 origin (2-norm), 2 ruiz (sqrt of inf-norm), 3 qp (sqrt(min nonzero * max))
// =========================================================================================================
__device__ __forceinline__ double clamp_scale(double v, double lo, double hi) { return v < lo ? 1.0 : (v > hi ? hi : v); }

__global__ void k_eq_cols(int kind, const int* cptr, double* val, int n, double* E_acc, double lo, double hi) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const int c1 = cptr[j], c2 = cptr[j + 1];
    double acc = 0.0;
    for (int k = c1; k < c2; ++k) {
        const double a = fabs(val[k]);
        if (kind == 0) acc = __dadd_rn(acc, a);
        else if (kind == 1) acc = __dadd_rn(acc, __dmul_rn(a, a));
        else acc = fmax(acc, a);
    }
    double e;
    if (kind == 3) {
        double mn = acc;
        for (int k = c1; k < c2; ++k) {
            const double a = fabs(val[k]);
            if (a <= mn && a > 0) mn = a;
        }
        e = __dmul_rn(sqrt(mn), sqrt(acc));
    } else {
        e = sqrt(acc);
    }
    e = clamp_scale(e, lo, hi);
    const double inv = 1.0 / e;
    for (int k = c1; k < c2; ++k) val[k] = __dmul_rn(val[k], inv);
    E_acc[j] = __dmul_rn(E_acc[j], e);
}

// row statistic in CSR order (ascending column = the reference's accumulation order), one thread per row
__global__ void k_eq_rows(int kind, const int* rptr, const int* perm, const double* val, int m, double* Dt, double* D_acc,
                          double lo, double hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int r1 = rptr[i], r2 = rptr[i + 1];
    double d = 0.0;
    for (int q = r1; q < r2; ++q) {
        const double a = fabs(val[perm[q]]);
        if (kind == 0) d = __dadd_rn(d, a);
        else if (kind == 1) d = __dadd_rn(d, __dmul_rn(a, a));
        else if (a >= d) d = a;
    }
    double v;
    if (kind == 3) {
        double mn = d;
        for (int q = r1; q < r2; ++q) {
            const double a = fabs(val[perm[q]]);
            if (a <= mn && a > 0) mn = a;
        }
        v = sqrt(__dmul_rn(d, mn));
    } else {
        v = sqrt(d);
    }
    v = clamp_scale(v, lo, hi);
    Dt[i] = v;
    D_acc[i] = __dmul_rn(D_acc[i], v);
}

__global__ void k_eq_apply_rows(const int* ridx, double* val, long nnz, const double* Dt) {
    const long k = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nnz) val[k] = val[k] / Dt[ridx[k]];
}

__global__ void k_fill(double* a, long n, double v) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void k_scale_all(double* a, long n, double f) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = __dmul_rn(a[i], f);
}

// per-row / per-column 2-norms of the scaled matrix divided by the count (common.c:535-557); summed on the host
__global__ void k_eq_row_norms(const int* rptr, const int* perm, const double* val, int m, double* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    double s = 0.0;
    for (int q = rptr[i]; q < rptr[i + 1]; ++q) { const double a = val[perm[q]]; s = __dadd_rn(s, __dmul_rn(a, a)); }
    out[i] = sqrt(s) / m;
}
__global__ void k_eq_col_norms(const int* cptr, const double* val, int n, double* out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    double s = 0.0;
    for (int k = cptr[j]; k < cptr[j + 1]; ++k) s = __dadd_rn(s, __dmul_rn(val[k], val[k]));
    out[j] = sqrt(s) / n;
}

__global__ void k_gather_perm(const double* src, const int* perm, double* dst, long nnz) {
    const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < nnz) dst[q] = src[perm[q]];
}

// Locality ordering (sjds_host.h: locality_order): the engine works in a permuted index space; vectors are permuted
// whenever they cross the C ABI.  n2o[i] = caller's index of the engine's entry i.
__global__ void k_perm_in(const double* src_old, const int* n2o, double* dst_new, int len, int valid) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) {
        const int o = n2o[i];
        if (o < valid) dst_new[i] = src_old[o];
    }
}
__global__ void k_perm_out(const double* src_new, const int* n2o, double* dst_old, int len) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < len) dst_old[n2o[i]] = src_new[i];
}

// =========================================================================================================
// Host side of the engine
// =========================================================================================================
#include "spmv_host.h"
#include "order_host.h"

struct ABIPGPU_LP {
    int device = 0;
    int m = 0, n = 0, l = 0;
    long nnz = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_solve0 = nullptr, ev_solve1 = nullptr;
    int num_sms = 0;
    int grid = 0, grid_mu = 0;  // every SpMV-bearing kernel uses the same persistent grid (the plan is per warp)
    // matrix
    int *A_ptr = nullptr, *A_idx = nullptr, *AT_ptr = nullptr, *AT_idx = nullptr, *A_wc = nullptr, *AT_wc = nullptr;
    unsigned char* arena = nullptr;              // all matrix / plan arrays live in this one allocation
    unsigned char* plan_arena = nullptr;         // plan arrays after the measured balance (tune_balance)
    double tune_ms = 0;                          // host time of the measured balance
    double tune_first_us[2] = {0, 0}, tune_best_us[2] = {0, 0};  // slowest CTA per pass (A', A): structural plan / chosen plan
    int tune_rounds = 0, tune_best_round = 0;
    int *A_cl = nullptr, *AT_cl = nullptr;       // long-row tables
    int4 *A_lr = nullptr, *AT_lr = nullptr;
    double *A_lp = nullptr, *AT_lp = nullptr;
    size_t smem = kSmemBytes;  // dynamic shared memory of the persistent kernels
    int4 *A_chunk = nullptr, *AT_chunk = nullptr;
    double *A_val = nullptr, *AT_val = nullptr;
    // everything else lives in one slab
    double* slab = nullptr;
    size_t slab_doubles = 0;
    double* vec[32] = {nullptr};
    long vec_len[32] = {0};
    double *dM = nullptr, *dD = nullptr, *dE = nullptr, *db = nullptr, *dc = nullptr;
    double *p = nullptr, *r = nullptr, *Gp = nullptr, *tmp = nullptr, *partials = nullptr, *dsc = nullptr;
    double* dphase = nullptr;
    double *xin = nullptr, *yout = nullptr;  // staging for host-pointer plugin calls [l] each
    double *xin2 = nullptr, *yout2 = nullptr;  // same, engine order (permuted engines only)
    bool permuted = false;                   // locality ordering active: engine index space != caller's
    int *d_rn2o = nullptr, *d_cn2o = nullptr, *d_pl = nullptr;  // new -> old maps: rows [m], columns [n], whole l-space
    double order_ms = 0;
    int resident_bytes = 0;               // batch engines: shared memory that keeps A and A' resident in k_batch (0: too large)
    double setup_ms[8] = {0};             // host set-up laps: transpose, ordering, permuted CSR, plans, upload, scaling, rest
    double* hsc = nullptr;                   // pinned host scalar block
    bool have_scaling = false;
    LpCtx ctx;
    ABIPSettings stgs;
    ABIPGpuStats stats;
    double B_A = 0, B_AT = 0;  // algorithmic bytes of one SpMV pass
    char desc[1024];
    bool restart_synced = false;
    struct BatchExec* batch = nullptr;  // lock-step batch executor this engine belongs to
    bool own_stream = true;             // batch engines borrow the stream of their worker thread
    bool dirty = false;                 // batch engines: asynchronous work was queued on the stream since the last step
    bool need_g = false;                // batch engines: g = K^-1 h is computed inside the next batched step (BatchItem::solve_g)
    int n_pend = 0;                     // batch engines: deferred vector operations (PreOp), executed by the next step
    PreOp pend[kMaxPre];
    // multi-GPU (column-block partition)
    int dist_G = 1, dist_rank = 0;
    long n_global = 0;
    unsigned char* comm_buf = nullptr;      // [vec 2*m_pad doubles | scal 2*kCommScalars doubles | flags kMaxRanks u64 | seq | err]
    size_t comm_bytes = 0;
    void* peer_bufs[kMaxRanks] = {nullptr};
};

static int coop_grid(const void* kernel, int num_sms, size_t smem, int* out) {
    int occ = 0;
    if (smem > 0) CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, kBlock, smem));
    if (occ < 1) {
        fprintf(stderr, "[abip_gpu] kernel cannot be made resident\n");
        return -1;
    }
    const int cap = env_int("ABIP_GPU_BLOCKS_PER_SM", ABIP_MIN_BLOCKS_PER_SM);
    *out = num_sms * std::min(occ, std::max(cap, 1));
    return 0;
}

template <class... Args>
static int launch_coop(abipgpu_lp* e, const void* kernel, int grid, size_t smem, Args... args) {
    void* argv[] = {(void*)&args...};
    // a one-CTA engine never reaches a grid barrier (grid_sync): an ordinary launch lets the engines of a batch run
    // concurrently (cooperative launches of different streams were observed to serialise)
    if (grid == 1) CK(cudaLaunchKernel(kernel, dim3(1), dim3(kBlock), argv, smem, e->stream));
    else CK(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kBlock), argv, smem, e->stream));
    e->stats.n_kernel_launches++;
    return 0;
}

static int read_sc(abipgpu_lp* e, abip_float* sc) {
    CK(cudaMemcpyAsync(e->hsc, e->dsc, sizeof(double) * ABIPGPU_SC_COUNT, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += sizeof(double) * ABIPGPU_SC_COUNT;
    if (sc) memcpy(sc, e->hsc, sizeof(double) * ABIPGPU_SC_COUNT);
    if (e->dist_G > 1 && e->hsc[ABIPGPU_SC_COMM_ERR] != 0.0) {
        fprintf(stderr, "[abip_gpu] rank %d: a peer GPU did not answer (multi-GPU collective timed out)\n", e->dist_rank);
        return -1;
    }
    return 0;
}

template <class T>
static int upload(T** dst, const std::vector<T>& src, abipgpu_lp* e) {
    const size_t bytes = (src.size() + kPad) * sizeof(T);  // fixed-size staging windows over-read behind the arrays
    CK(dev_alloc((void**)dst, bytes, e->stream));
    CK(cudaMemsetAsync(*dst, 0, bytes, e->stream));
    if (!src.empty()) CK(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, e->stream));
    e->stats.h2d_bytes += src.size() * sizeof(T);
    return 0;
}

// pinned host copies of the scalar block, recycled between engines (cudaMallocHost / cudaFreeHost synchronise)
#include <mutex>
static std::mutex g_pinned_mu;
static std::vector<double*> g_pinned_free;
static double* pinned_acquire() {
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        if (!g_pinned_free.empty()) {
            double* p = g_pinned_free.back();
            g_pinned_free.pop_back();
            return p;
        }
    }
    double* p = nullptr;
    if (cudaMallocHost((void**)&p, sizeof(double) * ABIPGPU_SC_COUNT) != cudaSuccess) return nullptr;
    return p;
}
static void pinned_release(double* p) {
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    g_pinned_free.push_back(p);
}

// =========================================================================================================
// Lock-step batch executor (BASELINE.json configs[4]).  Every problem in flight has its own host thread running the
// unchanged host solver (lp_host.cpp); a blocking step (ADMM iteration, BB round, mu statistics) is not launched by
// the thread itself but handed to the executor, which launches ONE k_batch for all requests that are waiting and wakes
// the threads when it has finished.  All engines of a batch share the executor's stream, so their occasional
// asynchronous operations (memsets, copies, small kernels) stay ordered with the batched steps.
// =========================================================================================================
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <thread>
#include <semaphore.h>
struct BatchReq {
    abipgpu_lp* e;
    BatchItem item;
    bool done = false;
    int rc = 0;
    sem_t sem;                   // per-request wake-up: no contention on the executor's mutex when a batch completes
    std::chrono::steady_clock::time_point t_submit;
    BatchReq() { sem_init(&sem, 0, 0); }
    ~BatchReq() { sem_destroy(&sem); }
};
constexpr int kResidentMax = 232448 - 4096;  // 227 KB per CTA on sm_100 minus the static shared memory of k_batch
struct BatchExec {
    int device = 0;
    int cap = 0;
    int wait_us = 200;
    bool use_resident = true;  // shared-memory-resident matrices when every problem of a launch fits (k_batch)
    // Launcher slots: each slot is a host thread with its own stream and its own host-mapped item / result buffers.  A
    // slot takes whatever requests are pending (after a short accumulation window), launches ONE k_batch for them,
    // synchronises its stream and wakes the owners -- while other slots launch the requests that arrive meanwhile.
    // With the device-resident loops a step lasts milliseconds and the steps of different problems differ widely (an
    // inner loop of 48 iterations, a BB search of 3 to 20 rounds): strict lock-step (one launch at a time, everybody
    // waits for the slowest item) left most CTAs idle (254 LP/s at cfg5 against 331 with one launch per iteration).
    struct Slot {
        cudaStream_t stream = nullptr;
        BatchItem* items = nullptr;  // host-mapped pinned memory, read by the kernel
        double* sc_out = nullptr;    // host-mapped pinned memory, written by the kernel
        std::thread th;
    };
    std::vector<Slot> slots;
    std::mutex mu;
    std::condition_variable cv_req;
    std::vector<BatchReq*> pending;
    int n_solving = 0;
    bool stop = false;
    std::atomic<long> n_launches{0}, n_items{0};
    double t_wait_ms = 0, t_kernel_ms = 0;
    bool early = true;          // release the owner of an item when its completion flag shows (ABIP_GPU_BATCH_EARLY=0: at launch end)
    std::atomic<long> seq{0};   // launch sequence numbers (completion flags)
    // statistics (ABIP_GPU_BATCH_VERBOSE): time of the launches, time of the items on their SMs, queueing of the requests
    std::mutex stat_mu;
    double st_launch_ms = 0, st_item_ms = 0, st_item_max_ms = 0, st_queue_ms = 0, st_release_lag_ms = 0;

    int start(int dev, int capacity) {
        device = dev;
        cap = capacity;
        wait_us = std::max(0, env_int("ABIP_GPU_BATCH_WAIT_US", 100));
        use_resident = env_int("ABIP_GPU_BATCH_RESIDENT", 1) != 0;
        early = env_int("ABIP_GPU_BATCH_EARLY", 1) != 0;
        // a slot is tied up for the length of its launch (a whole solve: ~70 ms at cfg5); with 6 slots a request waited
        // 15 - 35 ms for a free one (executor statistics, profiles/r02_batch.md)
        const int nslots = std::max(1, std::min(env_int("ABIP_GPU_BATCH_SLOTS", 16), 64));
        CK(cudaSetDevice(dev));
        CK(cudaFuncSetAttribute((const void*)k_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, kResidentMax));
        slots.resize(nslots);
        for (Slot& sl : slots) {
            CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
            CK(cudaHostAlloc((void**)&sl.items, sizeof(BatchItem) * cap, cudaHostAllocMapped));
            CK(cudaHostAlloc((void**)&sl.sc_out, sizeof(double) * ABIPGPU_SC_COUNT * cap, cudaHostAllocMapped));
        }
        for (int i = 0; i < nslots; ++i) slots[i].th = std::thread([this, i] { run(i); });
        return 0;
    }
    void run(int si) {
        cudaSetDevice(device);
        Slot& sl = slots[si];
        std::vector<BatchReq*> batch;
        std::unique_lock<std::mutex> lk(mu);
        for (;;) {
            cv_req.wait(lk, [&] { return stop || !pending.empty(); });
            if (stop && pending.empty()) break;
            if (wait_us > 0 && (int)pending.size() < std::min(cap, std::max(1, n_solving))) {
                // accumulation window: more requests usually arrive within a fraction of a millisecond
                lk.unlock();
                std::this_thread::sleep_for(std::chrono::microseconds(wait_us));
                lk.lock();
                if (pending.empty()) continue;  // another slot took them
            }
            batch.clear();
            const size_t take = std::min(pending.size(), (size_t)cap);
            batch.assign(pending.begin(), pending.begin() + take);
            pending.erase(pending.begin(), pending.begin() + take);
            lk.unlock();
            const int n = (int)batch.size();
            size_t smem = kSmemBytes;
            int resident = use_resident ? 1 : 0;
            for (int i = 0; i < n; ++i) {
                sl.items[i] = batch[i]->item;
                const int need = batch[i]->item.resident_bytes;
                if (need <= 0) resident = 0;
                else smem = std::max(smem, (size_t)need);
            }
            if (!resident) smem = kSmemBytes;
            const double my_seq = (double)(++seq);
            for (int i = 0; i < n; ++i) sl.sc_out[(size_t)i * ABIPGPU_SC_COUNT + ABIPGPU_SC_BATCH_DONE] = 0.0;
            const auto t_l0 = std::chrono::steady_clock::now();
            double queue_ms = 0, item_ms = 0, item_max = 0, lag_ms = 0;
            for (int i = 0; i < n; ++i) queue_ms += std::chrono::duration<double, std::milli>(t_l0 - batch[i]->t_submit).count();
            k_batch<<<n, kBlock, smem, sl.stream>>>(sl.items, sl.sc_out, resident, my_seq);
            cudaError_t err = cudaGetLastError();
            // release every item as soon as its CTA has written the completion flag (host-mapped memory); the items of one
            // launch differ by 2x in length and their owners have the next problem to set up
            auto release = [&](int i, bool ok) {
                const double* src = sl.sc_out + (size_t)i * ABIPGPU_SC_COUNT;
                const double ms = src[ABIPGPU_SC_BATCH_NS] * 1e-6;
                item_ms += ms;
                item_max = std::max(item_max, ms);
                memcpy(batch[i]->e->hsc, src, sizeof(double) * ABIPGPU_SC_COUNT);
                batch[i]->e->hsc[ABIPGPU_SC_BATCH_DONE] = 0.0;
                batch[i]->rc = ok ? 0 : -1;
                sem_post(&batch[i]->sem);  // the request may be gone right after this
                batch[i] = nullptr;
            };
            int remaining = n;
            unsigned polls = 0;
            if (err == cudaSuccess && early) {
                while (remaining > 0) {
                    for (int i = 0; i < n; ++i)
                        if (batch[i] && *(volatile double*)(sl.sc_out + (size_t)i * ABIPGPU_SC_COUNT + ABIPGPU_SC_BATCH_DONE) == my_seq) {
                            std::atomic_thread_fence(std::memory_order_acquire);
                            release(i, true);
                            --remaining;
                        }
                    if (remaining == 0) break;
                    if ((++polls & 31) == 0) {  // the stream itself only now and then (a driver call): failed launches
                        const cudaError_t q = cudaStreamQuery(sl.stream);
                        if (q != cudaErrorNotReady) {  // finished (flags are all set by now) or failed
                            if (q != cudaSuccess) err = q;
                            break;
                        }
                    }
                    // coarse while nothing can have finished yet (steps last milliseconds), fine afterwards
                    const double waited = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_l0).count();
                    std::this_thread::sleep_for(std::chrono::microseconds(waited < 2.0 ? 50 : (remaining == n ? 500 : 100)));
                }
            }
            if (err == cudaSuccess) err = cudaStreamSynchronize(sl.stream);
            if (err != cudaSuccess) fprintf(stderr, "[abip_gpu] batched step failed: %s\n", cudaGetErrorString(err));
            const auto t_l1 = std::chrono::steady_clock::now();
            for (int i = 0; i < n; ++i)
                if (batch[i]) {
                    const double ms = sl.sc_out[(size_t)i * ABIPGPU_SC_COUNT + ABIPGPU_SC_BATCH_NS] * 1e-6;
                    lag_ms += std::chrono::duration<double, std::milli>(t_l1 - t_l0).count() - ms;
                    release(i, err == cudaSuccess);
                }
            n_launches++;
            n_items += n;
            {
                std::lock_guard<std::mutex> sk(stat_mu);
                st_launch_ms += std::chrono::duration<double, std::milli>(t_l1 - t_l0).count();
                st_item_ms += item_ms;
                st_item_max_ms += item_max;
                st_queue_ms += queue_ms;
                st_release_lag_ms += lag_ms;
            }
            lk.lock();
        }
    }
    int submit(BatchReq* r) {
        r->t_submit = std::chrono::steady_clock::now();
        {
            std::lock_guard<std::mutex> lk(mu);
            pending.push_back(r);
        }
        cv_req.notify_one();
        while (sem_wait(&r->sem) != 0) {
        }
        return r->rc;
    }
    void solving(int delta) {
        std::lock_guard<std::mutex> lk(mu);
        n_solving += delta;
    }
    void finish() {
        {
            std::lock_guard<std::mutex> lk(mu);
            stop = true;
        }
        cv_req.notify_all();
        for (Slot& sl : slots)
            if (sl.th.joinable()) sl.th.join();
        cudaSetDevice(device);
        for (Slot& sl : slots) {
            if (sl.stream) {
                cudaStreamSynchronize(sl.stream);
                cudaStreamDestroy(sl.stream);
            }
            if (sl.items) cudaFreeHost(sl.items);
            if (sl.sc_out) cudaFreeHost(sl.sc_out);
        }
    }
};
static thread_local BatchExec* t_batch = nullptr;  // engines created by this thread join this executor
// Per-worker stream for everything that is not a batched step (set-up, equilibration, copies, small kernels): these
// must not queue behind the batched kernels of the other problems.  Ordering with the batched steps is by the host: a
// step is only submitted after the worker's stream has drained (BatchExec::submit_step), and the worker only
// continues after the executor has synchronised the batched launch.
static thread_local cudaStream_t t_worker_stream = nullptr;

extern "C" void* abipgpu_batch_begin(int device, int capacity) {
    BatchExec* b = new BatchExec();
    if (b->start(device, std::max(capacity, 1)) != 0) {
        delete b;
        return nullptr;
    }
    return b;
}
// One-time costs of a batch paid up front instead of by its first problems: the stream-ordered memory pool grows to the size
// the batch will need in ONE allocation (each of the ~400 first small allocations otherwise made the driver extend the pool:
// the first 512-problem batch of a process ran at 455 LP/s, the following ones at 1,290), and the pinned scalar blocks of
// all engines in flight come from one allocation.
extern "C" void abipgpu_batch_reserve(void* b, size_t device_bytes, int engines) {
    BatchExec* x = (BatchExec*)b;
    if (!x) return;
    if (cudaSetDevice(x->device) != cudaSuccess) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, x->device) == cudaSuccess) {
        unsigned long long keep = ~0ull, have = 0;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &have);
        if (have < device_bytes) {
            cudaStream_t st = nullptr;
            void* p = nullptr;
            if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess) {
                if (cudaMallocAsync(&p, device_bytes, st) == cudaSuccess) cudaFreeAsync(p, st);
                cudaStreamSynchronize(st);
                cudaStreamDestroy(st);
            }
            cudaGetLastError();
        }
    }
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    const int need = engines - (int)g_pinned_free.size();
    if (need > 0) {
        double* blk = nullptr;
        if (cudaMallocHost((void**)&blk, sizeof(double) * ABIPGPU_SC_COUNT * (size_t)need) == cudaSuccess)
            for (int i = 0; i < need; ++i) g_pinned_free.push_back(blk + (size_t)i * ABIPGPU_SC_COUNT);
        else
            cudaGetLastError();
    }
}
extern "C" void abipgpu_batch_attach(void* b) {
    t_batch = (BatchExec*)b;
    if (b && !t_worker_stream) {
        cudaSetDevice(t_batch->device);
        cudaStreamCreateWithFlags(&t_worker_stream, cudaStreamNonBlocking);
    } else if (!b && t_worker_stream) {
        cudaStreamSynchronize(t_worker_stream);
        cudaStreamDestroy(t_worker_stream);
        t_worker_stream = nullptr;
    }
}
extern "C" int abipgpu_batch_attached() { return t_batch != nullptr; }
extern "C" void abipgpu_batch_end(void* b, long* launches, long* items) {
    BatchExec* x = (BatchExec*)b;
    if (!x) return;
    x->finish();
    if (getenv("ABIP_GPU_BATCH_VERBOSE")) {
        const long nl = std::max(1L, x->n_launches.load()), ni = std::max(1L, x->n_items.load());
        printf("[abip_gpu] batch executor: %ld launches, %.1f items per launch; per launch %.1f ms (longest item %.1f ms); per item: "
               "%.1f ms on its SM, %.2f ms queued before the launch, %.2f ms between its end and its release\n",
               x->n_launches.load(), (double)ni / nl, x->st_launch_ms / nl, x->st_item_max_ms / nl, x->st_item_ms / ni,
               x->st_queue_ms / ni, x->st_release_lag_ms / ni);
    }
    if (launches) *launches = x->n_launches.load();
    if (items) *items = x->n_items.load();
    delete x;
}
static int flush_pending(abipgpu_lp* e);
// record a deferred operation (batch engines); falls back to executing everything directly when the list is full
static int defer(abipgpu_lp* e, PreOp op) {
    if (e->n_pend == kMaxPre && flush_pending(e)) return -1;
    e->pend[e->n_pend++] = op;
    return 0;
}
static int batch_step(abipgpu_lp* e, BatchReq* r, abip_float* sc) {
    r->item.n_pre = e->n_pend;
    for (int q = 0; q < e->n_pend; ++q) r->item.pre[q] = e->pend[q];
    e->n_pend = 0;
    for (int id = 0; id <= 20; ++id) r->item.vec[id] = e->vec[id];
    r->item.resident_bytes = e->resident_bytes;
    r->item.solve_g = e->need_g ? 1 : 0;
    if (e->dirty) {  // copies / memsets / small kernels queued by this thread must have finished
        CK(cudaStreamSynchronize(e->stream));
        e->dirty = false;
    }
    if (e->batch->submit(r)) return -1;
    if (e->need_g) {  // the step computed g and g_th first
        e->ctx.g_th = e->hsc[ABIPGPU_SC_VEC_NORM2];
        e->need_g = false;
    }
    if (sc) memcpy(sc, e->hsc, sizeof(double) * ABIPGPU_SC_COUNT);
    return 0;
}
void abipgpu_lp_batch_solving(abipgpu_lp* e, int delta) {
    if (e && e->batch) e->batch->solving(delta);
}

static thread_local int t_grid_request = 0;  // CTAs per engine (batch mode); 0 = whole device
extern "C" void abipgpu_lp_request_grid(int ctas) { t_grid_request = ctas; }
static thread_local int t_order_request = 1;  // 0: engines created by this thread keep the caller's row / column order
extern "C" void abipgpu_lp_request_order(int on) { t_order_request = on; }

struct ScaleOut {  // host outputs of the device-side equilibration
    double *D, *E, *mean_row, *mean_col;
};

// Device properties and the occupancy-derived persistent grids are the same for every engine of a process: query them
// once per device (cudaGetDeviceProperties alone costs more than a whole ADMM iteration of a small LP).
struct DevInfo {
    bool ok = false;
    int coop = 0, num_sms = 0;
    char name[256] = {0};
    int g_main = 0, g_mu = 0;  // grids for smem = kSmemBytes (no page cache)
};
static std::mutex g_devinfo_mu;
static DevInfo g_devinfo[64];


// CSR(A') is the caller's CSC; CSR(A) by counting sort (the reference's transpose(), indirect.c:81-139);
// perm[q] = position in the CSC arrays of entry q of CSR(A).  Every thread takes a range of columns: counts per row
// first, then fills its entries behind those of the threads before it, so that every row lists its columns in ascending
// order whatever the thread count.
// Equilibration sweeps on the device (kernels above) in the CALLER's order: d_val = CSC values (in: A, out: scaled A),
// dD [m], dE [n] accumulate the factors; Dt [m] and rn [max(m, n)] are scratch.  Bit-identical to abip_normalize_A.
static int run_equilibration(cudaStream_t st, int m, int n, long nnz, const int* d_atptr, const int* d_atidx, const int* d_aptr,
                             const int* d_perm, double* d_val, double* dD, double* dE, double* Dt, double* rn,
                             const ABIPSettings* stgs, double* mean_row, double* mean_col) {
    const double min_row = 1e-3 * sqrt((double)n), max_row = 1e3 * sqrt((double)n);
    const double min_col = 1e-3 * sqrt((double)m), max_col = 1e3 * sqrt((double)m);
    const unsigned gm = (unsigned)((m + 255) / 256), gn = (unsigned)((n + 255) / 256), gz = (unsigned)((nnz + 255) / 256);
    k_fill<<<gm, 256, 0, st>>>(dD, m, 1.0);
    k_fill<<<gn, 256, 0, st>>>(dE, n, 1.0);
    auto sweep = [&](int kind) {
        k_eq_cols<<<gn, 256, 0, st>>>(kind, d_atptr, d_val, n, dE, min_col, max_col);
        k_eq_rows<<<gm, 256, 0, st>>>(kind, d_aptr, d_perm, d_val, m, Dt, dD, min_row, max_row);
        k_eq_apply_rows<<<gz, 256, 0, st>>>(d_atidx, d_val, nnz, Dt);
    };
    if (stgs->pc_ruiz_rescale) sweep(0);
    if (stgs->origin_rescale) sweep(1);
    if (stgs->pc_ruiz_rescale)
        for (abip_int it = 0; it < stgs->ruiz_iter; ++it) sweep(2);
    if (stgs->qp_rescale) sweep(3);
    // mean row / column norms (summed on the host in index order, like common.c:541-557)
    std::vector<double> hn(std::max(m, n));
    k_eq_row_norms<<<gm, 256, 0, st>>>(d_aptr, d_perm, d_val, m, rn);
    CK(cudaMemcpyAsync(hn.data(), rn, sizeof(double) * m, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double mr = 0.0;
    for (int i = 0; i < m; ++i) mr += hn[i];
    k_eq_col_norms<<<gn, 256, 0, st>>>(d_atptr, d_val, n, rn);
    CK(cudaMemcpyAsync(hn.data(), rn, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    double mc = 0.0;
    for (int j = 0; j < n; ++j) mc += hn[j];
    *mean_row = mr;
    *mean_col = mc;
    if (stgs->scale != 1) k_scale_all<<<gz, 256, 0, st>>>(d_val, nnz, stgs->scale);
    CK(cudaGetLastError());
    return 0;
}

static int transpose_csc(abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai, int host_threads, std::vector<int>* at_ptr_,
                         std::vector<int>* at_idx_, std::vector<int>* a_ptr_, std::vector<int>* a_idx_, std::vector<int>* perm_) {
    const long nnz = Ap[n];
    std::vector<int>&at_ptr = *at_ptr_, &at_idx = *at_idx_, &a_ptr = *a_ptr_, &a_idx = *a_idx_, &perm = *perm_;
    at_ptr.resize(n + 1);
    at_idx.resize(nnz);
    a_ptr.assign(m + 1, 0);
    a_idx.resize(nnz);
    perm.resize(nnz);
    auto par = [&](long cnt, auto fn) { parallel_for(cnt, host_threads, fn); };
    for (long j = 0; j <= n; ++j) at_ptr[j] = (int)Ap[j];
    {
        std::vector<std::vector<int>> cnt(host_threads);
        std::vector<int> bad(host_threads, 0);
        par(n, [&](long j0, long j1, int t) {
            std::vector<int>& c = cnt[t];
            c.assign(m, 0);
            for (long k = Ap[j0]; k < Ap[j1]; ++k) {
                const long r = Ai[k];
                if (r < 0 || r >= m) { bad[t] = 1; return; }
                at_idx[k] = (int)r;
                c[r]++;
            }
        });
        for (int t = 0; t < host_threads; ++t)
            if (bad[t]) {
                fprintf(stderr, "[abip_gpu] row index out of range\n");
                return -1;
            }
        for (long i = 0; i < m; ++i) {
            int tot = 0;
            for (int t = 0; t < host_threads; ++t) {
                if (cnt[t].empty()) continue;  // (fewer ranges than threads)
                const int c = cnt[t][i];
                cnt[t][i] = tot;
                tot += c;
            }
            a_ptr[i + 1] = a_ptr[i] + tot;
        }
        par(n, [&](long j0, long j1, int t) {
            std::vector<int>& off = cnt[t];
            for (long j = j0; j < j1; ++j)
                for (long k = Ap[j]; k < Ap[j + 1]; ++k) {
                    const int r = at_idx[k];
                    const int q = a_ptr[r] + off[r]++;
                    a_idx[q] = (int)j;
                    perm[q] = (int)k;
                }
        });
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Measured balance of the persistent grid.  A phase of the PCG loop lasts as long as its slowest CTA; the structural cost
// model of build_spmv_plan (nonzeros + rows + distinct gathered lines) leaves the CTA times of a pass spread by +-25 % at
// cfg2 (profiles/r02_phase_times.txt: per-CTA mean busy time 20-32 us for A', 24-37 us for A) -- gather cost depends on what
// the neighbouring CTAs of the SM and of the L2 slice do, which no static model sees.  So the engine measures: k_tune runs
// the two passes on the real matrices, the model cost of the rows of every CTA is rescaled by (its time / mean time)^damp,
// the row ranges are cut again on the rescaled costs, and the best plan seen (smallest sum of the slowest-CTA times of
// the two passes) is kept.  Only the plan tables change (a few hundred KB); row sums do not depend on the plan, the order
// of the reduced scalars does (the same way it depends on the grid size), so a solve is reproducible for a given plan.
// ABIP_GPU_TUNE=0 keeps the structural plan; ABIP_GPU_TUNE_ROUNDS / _REPS / _DAMP / _MIN_NNZ are tuning knobs.
// ---------------------------------------------------------------------------------------------------------
static int upload_plans(abipgpu_lp* e, const SpmvPlan& pa, const SpmvPlan& pat) {
    size_t bytes = 0;
    auto put = [&](size_t b, size_t elem) {
        const size_t off = (bytes + 255) & ~(size_t)255;
        bytes = off + b + (size_t)kPad * elem;
        return off;
    };
    struct Offs { size_t wc, ch, cl, lr, lp; } o[2];
    const SpmvPlan* P[2] = {&pa, &pat};
    for (int k = 0; k < 2; ++k) {
        o[k].wc = put(P[k]->warp_chunk.size() * sizeof(int), sizeof(int));
        o[k].ch = put(P[k]->chunk.size() * sizeof(int4), sizeof(int4));
        o[k].cl = o[k].lr = o[k].lp = 0;
        if (P[k]->n_long) {
            o[k].cl = put(P[k]->cta_long.size() * sizeof(int), sizeof(int));
            o[k].lr = put(P[k]->long_rows.size() * sizeof(int4), sizeof(int4));
            o[k].lp = put((size_t)P[k]->n_pieces * sizeof(double), sizeof(double));
        }
    }
    std::vector<unsigned char> stage(bytes, 0);
    for (int k = 0; k < 2; ++k) {
        memcpy(stage.data() + o[k].wc, P[k]->warp_chunk.data(), P[k]->warp_chunk.size() * sizeof(int));
        memcpy(stage.data() + o[k].ch, P[k]->chunk.data(), P[k]->chunk.size() * sizeof(int4));
        if (P[k]->n_long) {
            memcpy(stage.data() + o[k].cl, P[k]->cta_long.data(), P[k]->cta_long.size() * sizeof(int));
            memcpy(stage.data() + o[k].lr, P[k]->long_rows.data(), P[k]->long_rows.size() * sizeof(int4));
        }
    }
    unsigned char* buf = nullptr;
    CK(dev_alloc((void**)&buf, bytes, e->stream));
    CK(cudaMemcpyAsync(buf, stage.data(), bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));  // also: every launch that used the previous tables has finished
    e->stats.h2d_bytes += (double)bytes;
    if (e->plan_arena) dev_free(e->plan_arena, e->stream);
    e->plan_arena = buf;
    e->A_wc = (int*)(buf + o[0].wc); e->A_chunk = (int4*)(buf + o[0].ch);
    e->AT_wc = (int*)(buf + o[1].wc); e->AT_chunk = (int4*)(buf + o[1].ch);
    e->A_cl = pa.n_long ? (int*)(buf + o[0].cl) : nullptr;
    e->A_lr = pa.n_long ? (int4*)(buf + o[0].lr) : nullptr;
    e->A_lp = pa.n_long ? (double*)(buf + o[0].lp) : nullptr;
    e->AT_cl = pat.n_long ? (int*)(buf + o[1].cl) : nullptr;
    e->AT_lr = pat.n_long ? (int4*)(buf + o[1].lr) : nullptr;
    e->AT_lp = pat.n_long ? (double*)(buf + o[1].lp) : nullptr;
    Csr& A = e->ctx.A;
    A.warp_chunk = e->A_wc; A.chunk = e->A_chunk; A.lanes_log2 = pa.lanes_log2;
    A.cta_long = e->A_cl; A.long_rows = e->A_lr; A.long_part = e->A_lp;
    Csr& AT = e->ctx.AT;
    AT.warp_chunk = e->AT_wc; AT.chunk = e->AT_chunk; AT.lanes_log2 = pat.lanes_log2;
    AT.cta_long = e->AT_cl; AT.long_rows = e->AT_lr; AT.long_part = e->AT_lp;
    return 0;
}

static int tune_balance(abipgpu_lp* e, const std::vector<int>& a_ptr, const std::vector<int>& at_ptr, SpmvPlan* planA,
                        SpmvPlan* planAT, int W, bool threads) {
    const auto t_begin = std::chrono::steady_clock::now();
    const int G = e->grid;
    const int rounds = std::max(0, std::min(env_int("ABIP_GPU_TUNE_ROUNDS", 3), 8));
    const int reps = std::max(2, std::min(env_int("ABIP_GPU_TUNE_REPS", 4), 64));
    const char* denv = getenv("ABIP_GPU_TUNE_DAMP");
    const double damp = (denv && *denv) ? atof(denv) : 0.8;
    const std::vector<int>* ptr[2] = {&a_ptr, &at_ptr};
    const int nrows[2] = {e->m, e->n};
    const char* lanes_env[2] = {"ABIP_GPU_LANES_A", "ABIP_GPU_LANES_AT"};
    SpmvPlan cur[2], best[2];
    std::vector<double> cost[2];
    cost[0].swap(planA->row_cost);
    cost[1].swap(planAT->row_cost);
    if ((int)cost[0].size() != e->m || (int)cost[1].size() != e->n) return 0;  // no model costs (one-CTA engines)
    cur[0] = *planA;
    cur[1] = *planAT;
    CK(cudaFuncSetAttribute((const void*)k_tune, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
    double* d_t = nullptr;
    CK(dev_alloc((void**)&d_t, sizeof(double) * 11 * G, e->stream));
    std::vector<double> t(11 * (size_t)G);
    double best_score = 1e300;
    int best_round = -1, uploaded_round = 0;
    int cl = 0;
    cudaDeviceGetAttribute(&cl, cudaDevAttrClockRate, e->device);  // kHz
    const double us_per_cycle = cl > 0 ? 1e3 / (double)cl : 1.0 / 1965.0;
    for (int round = 0; round <= rounds; ++round) {
        if (launch_coop(e, (const void*)k_tune, G, e->smem, e->ctx, d_t, reps)) return -1;
        CK(cudaMemcpyAsync(t.data(), d_t, sizeof(double) * 11 * G, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaStreamSynchronize(e->stream));
        double mx[2] = {0, 0}, mean[2] = {0, 0};
        for (int k = 0; k < 2; ++k) {
            const double* tk = t.data() + (k == 0 ? G : 0);  // t = [A' | A | sink]; k = 0 is A
            for (int b = 0; b < G; ++b) {
                mx[k] = std::max(mx[k], tk[b]);
                mean[k] += tk[b] / G;
            }
        }
        const double score = mx[0] + mx[1];
        const double to_us = us_per_cycle / (reps - 1);
        if (round == 0) {
            e->tune_first_us[0] = mx[1] * to_us;
            e->tune_first_us[1] = mx[0] * to_us;
        }
        if (env_int("ABIP_GPU_TUNE_VERBOSE", 0))
            fprintf(stderr, "[abip_gpu] balance round %d: slowest CTA A' %.1f us (mean %.1f), A %.1f us (mean %.1f)\n", round,
                    mx[1] * to_us, mean[1] * to_us, mx[0] * to_us, mean[0] * to_us);
        if (env_int("ABIP_GPU_TUNE_VERBOSE", 0) > 1) {
            static const char* nm[8] = {"A' pass", "barrier 1 wait", "A pass", "block_store", "barrier 2 wait", "finish", "update", "barrier 3 wait"};
            double tot = 0;
            for (int k = 0; k < 8; ++k) {
                const double* sk = t.data() + (size_t)(3 + k) * G;
                double mn = 1e300, mxs = 0, av = 0;
                for (int b = 0; b < G; ++b) { mn = std::min(mn, sk[b]); mxs = std::max(mxs, sk[b]); av += sk[b] / G; }
                fprintf(stderr, "[abip_gpu]   section %-15s us per iteration: mean %6.2f  min %6.2f  max %6.2f\n", nm[k], av * to_us, mn * to_us, mxs * to_us);
                tot += av * to_us;
            }
            fprintf(stderr, "[abip_gpu]   sum of the means %.2f us per iteration\n", tot);
        }
        if (score < best_score) {
            best_score = score;
            best_round = round;
            best[0] = cur[0];
            best[1] = cur[1];
            e->tune_best_us[0] = mx[1] * to_us;
            e->tune_best_us[1] = mx[0] * to_us;
        }
        if (round == rounds) break;
        auto replan = [&](int k) {
            const double* tk = t.data() + (k == 0 ? G : 0);
            if (!(mean[k] > 0)) return;
            std::vector<double>& c = cost[k];
            const std::vector<int>& cr = cur[k].cta_row;
            for (int b = 0; b < G; ++b) {
                double f = pow(std::max(tk[b], 1.0) / mean[k], damp);
                f = std::min(2.0, std::max(0.5, f));
                for (int r = cr[b]; r < cr[b + 1]; ++r) c[r] *= f;
            }
            SpmvPlan np;
            build_spmv_plan(*ptr[k], nrows[k], W, lanes_env[k], &np, nullptr, &c);
            cur[k] = std::move(np);
        };
        if (threads) {
            std::thread tp([&] { replan(0); });
            replan(1);
            tp.join();
        } else {
            replan(0);
            replan(1);
        }
        if (upload_plans(e, cur[0], cur[1])) return -1;
        uploaded_round = round + 1;
    }
    if (best_round != uploaded_round && upload_plans(e, best[0], best[1])) return -1;
    dev_free(d_t, e->stream);
    // the PCG workspace stays zero (alpha = 0 in k_tune), Gp and tmp are scratch
    *planA = std::move(best[0]);
    *planAT = std::move(best[1]);
    e->tune_rounds = rounds;
    e->tune_best_round = best_round;
    e->tune_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return 0;
}

static int create_impl(abipgpu_lp* e, abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai,
                       const abip_float* Ax, const ABIPSettings* stgs, int device, const ScaleOut* scale_out = nullptr) {
    const long nnz = Ap[n];
    if (m <= 0 || n <= 0 || nnz <= 0 || nnz >= 2147483647L || (long)m + n + 1 >= 2147483647L) {
        fprintf(stderr, "[abip_gpu] unsupported size m=%ld n=%ld nnz=%ld (int32 device indices)\n", (long)m,
                (long)n, nnz);
        return -1;
    }
    e->device = device;
    e->m = (int)m;
    e->n = (int)n;
    e->l = (int)(m + n + 1);
    e->nnz = nnz;
    e->stgs = *stgs;
    memset(&e->stats, 0, sizeof(e->stats));
    CK(cudaSetDevice(device));
    if (device < 0 || device >= 64) return -1;
    DevInfo& di = g_devinfo[device];
    {
        std::lock_guard<std::mutex> lk(g_devinfo_mu);
        if (!di.ok) {
            cudaDeviceProp p0;
            CK(cudaGetDeviceProperties(&p0, device));
            di.coop = p0.cooperativeLaunch;
            di.num_sms = p0.multiProcessorCount;
            snprintf(di.name, sizeof(di.name), "%s", p0.name);
            di.ok = true;
        }
    }
    struct { int cooperativeLaunch, multiProcessorCount; const char* name; } prop = {di.coop, di.num_sms, di.name};
    if (!prop.cooperativeLaunch) {
        fprintf(stderr, "[abip_gpu] device lacks cooperative launch\n");
        return -1;
    }
    e->num_sms = prop.multiProcessorCount;
    if (t_batch && t_worker_stream) {  // lock-step batch: one CTA per problem, the worker thread's stream, no events
        e->batch = t_batch;
        {
            const size_t need = resident_bytes_for(m, n, nnz);
            e->resident_bytes = (m < 65536 && n < 65536 && need <= (size_t)kResidentMax) ? (int)need : 0;
        }
        e->stream = t_worker_stream;
        e->own_stream = false;
        e->dirty = true;
    } else {
        CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
        CK(cudaEventCreate(&e->ev0));
        CK(cudaEventCreate(&e->ev1));
        CK(cudaEventCreate(&e->ev_solve0));
        CK(cudaEventCreate(&e->ev_solve1));
    }

    auto lap_t = std::chrono::steady_clock::now();
    auto lap = [&](int slot) {
        const auto t = std::chrono::steady_clock::now();
        e->setup_ms[slot] += std::chrono::duration<double, std::milli>(t - lap_t).count();
        lap_t = t;
    };
    const int host_threads = (nnz < 200000 || t_batch) ? 1 : std::max(1, std::min(env_int("ABIP_GPU_HOST_THREADS", 8), (int)std::thread::hardware_concurrency()));
    auto par = [&](long cnt, auto fn) { parallel_for(cnt, host_threads, fn); };
    std::vector<int> at_ptr, at_idx, a_ptr, a_idx, perm;
    if (transpose_csc(m, n, Ap, Ai, host_threads, &at_ptr, &at_idx, &a_ptr, &a_idx, &perm)) return -1;
    // Locality ordering (sjds_host.h): whole-device engines work in a permuted index space in which structurally
    // identical rows / columns are neighbours, so the 32 lanes of a gather touch a few lines instead of 32.
    // e_* : the engine's matrices; e_a_src / e_at_src: position of every entry in the caller's CSC arrays.
    lap(0);
    std::vector<int> row_n2o, col_n2o;
    bool reorder = env_int("ABIP_GPU_REORDER", 1) != 0 && t_order_request != 0 && !t_batch && t_grid_request == 0;
    if (reorder) {
        const auto t_o0 = std::chrono::steady_clock::now();
        sjds::locality_order((int)m, (int)n, a_ptr, a_idx, at_ptr, at_idx, sjds::kLongRow, &row_n2o, &col_n2o, par);
        bool ident = true;
        for (long i = 0; i < m && ident; ++i) ident = row_n2o[i] == i;
        for (long j = 0; j < n && ident; ++j) ident = col_n2o[j] == j;
        if (ident) reorder = false;
        e->order_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_o0).count();
    }
    e->permuted = reorder;
    lap(1);
    std::vector<int> e_a_ptr, e_a_idx, e_a_src, e_at_ptr, e_at_idx, e_at_src;
    if (reorder) {
        std::vector<int> row_o2n(m), col_o2n(n);
        for (long i = 0; i < m; ++i) row_o2n[row_n2o[i]] = (int)i;
        for (long j = 0; j < n; ++j) col_o2n[col_n2o[j]] = (int)j;
        e_a_ptr.assign(m + 1, 0);
        e_a_idx.resize(nnz);
        e_a_src.resize(nnz);
        for (long i = 0; i < m; ++i) e_a_ptr[i + 1] = e_a_ptr[i] + (a_ptr[row_n2o[i] + 1] - a_ptr[row_n2o[i]]);
        par(m, [&](long i0, long i1, int) {
            for (long i = i0; i < i1; ++i) {
                int q = e_a_ptr[i];
                for (int k = a_ptr[row_n2o[i]]; k < a_ptr[row_n2o[i] + 1]; ++k, ++q) {
                    e_a_idx[q] = col_o2n[a_idx[k]];
                    e_a_src[q] = perm[k];
                }
            }
        });
        e_at_ptr.assign(n + 1, 0);
        e_at_idx.resize(nnz);
        e_at_src.resize(nnz);
        for (long j = 0; j < n; ++j) e_at_ptr[j + 1] = e_at_ptr[j] + (at_ptr[col_n2o[j] + 1] - at_ptr[col_n2o[j]]);
        par(n, [&](long j0, long j1, int) {
            for (long j = j0; j < j1; ++j) {
                int q = e_at_ptr[j];
                for (int k = at_ptr[col_n2o[j]]; k < at_ptr[col_n2o[j] + 1]; ++k, ++q) {
                    e_at_idx[q] = row_o2n[at_idx[k]];
                    e_at_src[q] = k;
                }
            }
        });
    } else {
        e_a_ptr = a_ptr;
        e_a_idx = a_idx;
        e_a_src = perm;
        e_at_ptr = at_ptr;
        e_at_idx = at_idx;
    }
    std::vector<double> a_val(scale_out ? 0 : nnz), at_val(scale_out ? 0 : nnz);
    if (!scale_out) {
        par(nnz, [&](long q0, long q1, int) {
            for (long q = q0; q < q1; ++q) a_val[q] = Ax[e_a_src[q]];
            if (reorder) for (long q = q0; q < q1; ++q) at_val[q] = Ax[e_at_src[q]];
            else memcpy(at_val.data() + q0, Ax + q0, sizeof(double) * (q1 - q0));
        });
    }
    lap(2);
    // persistent grid: a multiple of the SM count, common to all SpMV-bearing kernels
    {
        int g1, g2, g3, g4;
        int h1, h2, h3;
        e->smem = kSmemBytes;
        bool cached = false;
        {
            std::lock_guard<std::mutex> lk(g_devinfo_mu);
            if (di.g_main > 0) {
                g1 = g2 = g3 = h1 = h2 = h3 = di.g_main;
                g4 = di.g_mu;
                cached = true;
            }
        }
        if (cached) {
        } else if (coop_grid((const void*)k_admm_iter<false>, e->num_sms, e->smem, &g1) ||
            coop_grid((const void*)k_bb_round<false>, e->num_sms, e->smem, &g2) ||
            coop_grid((const void*)k_solve_vec<false>, e->num_sms, e->smem, &g3) ||
            coop_grid((const void*)k_admm_iter<true>, e->num_sms, e->smem, &h1) ||
            coop_grid((const void*)k_bb_round<true>, e->num_sms, e->smem, &h2) ||
            coop_grid((const void*)k_solve_vec<true>, e->num_sms, e->smem, &h3) ||
            coop_grid((const void*)k_mu_stats, e->num_sms, 0, &g4))
            return -1;
        g1 = std::min(g1, h1); g2 = std::min(g2, h2); g3 = std::min(g3, h3);
        CK(cudaFuncSetAttribute((const void*)k_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem));
        e->grid = std::min(g1, std::min(g2, g3));
        e->grid_mu = g4;
        if (!cached) {
            std::lock_guard<std::mutex> lk(g_devinfo_mu);
            di.g_main = e->grid;
            di.g_mu = e->grid_mu;
        }
        if (e->batch) {
            e->grid = 1;
            e->grid_mu = 1;
        } else if (t_grid_request > 0) {  // batch mode: several small engines share the device
            e->grid = std::min(e->grid, t_grid_request);
            e->grid_mu = std::min(e->grid_mu, t_grid_request);
        }
    }
    const int W = e->grid * kWarps;
    // measured balance (tune_balance; opt-in: measured at cfg2 and found to change nothing, profiles/r02_spmv.md): whole-device
    // engines of at least ABIP_GPU_TUNE_MIN_NNZ nonzeros, contiguous row ranges
    const bool tune = env_int("ABIP_GPU_TUNE", 0) != 0 && !e->batch && t_grid_request == 0 && e->grid > 1 &&
                      nnz >= (long)env_int("ABIP_GPU_TUNE_MIN_NNZ", 1000000) && env_int("ABIP_GPU_PLAN_DEAL", 0) == 0;
    SpmvPlan planA, planAT;
    if (host_threads > 1) {
        std::thread tp([&] { build_spmv_plan(e_a_ptr, (int)m, W, "ABIP_GPU_LANES_A", &planA, e_a_idx.data(), nullptr, tune); });
        build_spmv_plan(e_at_ptr, (int)n, W, "ABIP_GPU_LANES_AT", &planAT, e_at_idx.data(), nullptr, tune);
        tp.join();
    } else {
        build_spmv_plan(e_a_ptr, (int)m, W, "ABIP_GPU_LANES_A", &planA, e_a_idx.data(), nullptr, tune);
        build_spmv_plan(e_at_ptr, (int)n, W, "ABIP_GPU_LANES_AT", &planAT, e_at_idx.data(), nullptr, tune);
    }
    lap(3);
    // One device arena for all matrix and plan arrays: packed on the host (zero padding of kPad elements behind every
    // array included) and uploaded with ONE allocation and ONE copy -- 14 arrays x (malloc + memset + copy) were a
    // third of the driver calls of an engine set-up, which is what limits a batch of small LPs.
    // (large problems skip the host staging copy: one memset of the arena, then one copy per array)
    std::vector<unsigned char> stage;
    struct Seg { size_t off; const void* src; size_t bytes; };
    std::vector<Seg> segs;
    size_t arena_bytes = 0;
    auto put = [&](const void* src, size_t bytes, size_t elem) -> size_t {
        const size_t off = (arena_bytes + 255) & ~(size_t)255;
        arena_bytes = off + bytes + (size_t)kPad * elem;
        if (src && bytes) segs.push_back(Seg{off, src, bytes});
        return off;
    };
    auto put_vec = [&](const auto& v) { return put(v.data(), v.size() * sizeof(v[0]), sizeof(v[0])); };
    const size_t o_aptr = put_vec(e_a_ptr), o_aidx = put_vec(e_a_idx);
    const size_t o_aval = scale_out ? put(nullptr, nnz * sizeof(double), sizeof(double)) : put_vec(a_val);
    const size_t o_atptr = put_vec(e_at_ptr), o_atidx = put_vec(e_at_idx);
    const size_t o_atval = scale_out ? put(nullptr, nnz * sizeof(double), sizeof(double)) : put_vec(at_val);
    const size_t o_awc = put_vec(planA.warp_chunk), o_atwc = put_vec(planAT.warp_chunk);
    const size_t o_ach = put_vec(planA.chunk), o_atch = put_vec(planAT.chunk);
    size_t o_acl = 0, o_alr = 0, o_alp = 0, o_atcl = 0, o_atlr = 0, o_atlp = 0;
    if (planA.n_long) {
        o_acl = put_vec(planA.cta_long);
        o_alr = put_vec(planA.long_rows);
        o_alp = put(nullptr, planA.n_pieces * sizeof(double), sizeof(double));
    }
    if (planAT.n_long) {
        o_atcl = put_vec(planAT.cta_long);
        o_atlr = put_vec(planAT.long_rows);
        o_atlp = put(nullptr, planAT.n_pieces * sizeof(double), sizeof(double));
    }
    CK(dev_alloc((void**)&e->arena, arena_bytes, e->stream));
    if (arena_bytes <= ((size_t)16 << 20)) {
        stage.assign(arena_bytes, 0);
        for (const Seg& sg : segs) memcpy(stage.data() + sg.off, sg.src, sg.bytes);
        CK(cudaMemcpyAsync(e->arena, stage.data(), arena_bytes, cudaMemcpyHostToDevice, e->stream));
        e->stats.h2d_bytes += (double)arena_bytes;
    } else {
        CK(cudaMemsetAsync(e->arena, 0, arena_bytes, e->stream));
        for (const Seg& sg : segs) {
            CK(cudaMemcpyAsync(e->arena + sg.off, sg.src, sg.bytes, cudaMemcpyHostToDevice, e->stream));
            e->stats.h2d_bytes += (double)sg.bytes;
        }
    }
    unsigned char* ab = e->arena;
    e->A_ptr = (int*)(ab + o_aptr); e->A_idx = (int*)(ab + o_aidx); e->A_val = (double*)(ab + o_aval);
    e->AT_ptr = (int*)(ab + o_atptr); e->AT_idx = (int*)(ab + o_atidx); e->AT_val = (double*)(ab + o_atval);
    e->A_wc = (int*)(ab + o_awc); e->AT_wc = (int*)(ab + o_atwc);
    e->A_chunk = (int4*)(ab + o_ach); e->AT_chunk = (int4*)(ab + o_atch);
    if (planA.n_long) { e->A_cl = (int*)(ab + o_acl); e->A_lr = (int4*)(ab + o_alr); e->A_lp = (double*)(ab + o_alp); }
    if (planAT.n_long) { e->AT_cl = (int*)(ab + o_atcl); e->AT_lr = (int4*)(ab + o_atlr); e->AT_lp = (double*)(ab + o_atlp); }
    const int gmax = std::max(e->grid, e->grid_mu);

    // one slab for all FP64 vectors, each 256-byte aligned
    const size_t L = ((size_t)e->l + 31) & ~(size_t)31;
    const size_t Mm = ((size_t)m + 31) & ~(size_t)31, Nn = ((size_t)n + 31) & ~(size_t)31;
    const size_t n_l_vecs = 21 + 2 + (reorder ? 2 : 0);  // ids 0..20 + xin + yout (+ xin2 + yout2)
    const size_t total = n_l_vecs * L + 7 * Mm + 3 * Nn + (size_t)2 * kMaxRed * gmax + 64 + ABIPGPU_SC_COUNT + 32 + 2 * (size_t)gmax * kWarps + gmax;
    CK(dev_alloc((void**)&e->slab, total * sizeof(double), e->stream));
    CK(cudaMemsetAsync(e->slab, 0, total * sizeof(double), e->stream));
    e->slab_doubles = total;
    double* q = e->slab;
    auto take = [&](size_t k) { double* r = q; q += k; return r; };
    for (int id = 0; id <= 20; ++id) {
        e->vec[id] = take(L);
        e->vec_len[id] = e->l;
    }
    e->vec_len[ABIPGPU_VEC_H] = e->vec_len[ABIPGPU_VEC_G] = m + n;
    e->xin = take(L);
    e->yout = take(L);
    if (reorder) {
        e->xin2 = take(L);
        e->yout2 = take(L);
        std::vector<int> pl(e->l);
        for (long i = 0; i < m; ++i) pl[i] = row_n2o[i];
        for (long j = 0; j < n; ++j) pl[m + j] = (int)m + col_n2o[j];
        pl[m + n] = (int)(m + n);
        if (upload(&e->d_pl, pl, e) || upload(&e->d_rn2o, row_n2o, e) || upload(&e->d_cn2o, col_n2o, e)) return -1;
        CK(cudaStreamSynchronize(e->stream));  // pl goes out of scope
    }
    e->dM = take(Mm);
    e->vec[ABIPGPU_VEC_M] = e->dM;  // overrides the l-sized slot: M has its own storage
    e->vec_len[ABIPGPU_VEC_M] = m;
    e->dD = take(Mm);
    e->db = take(Mm);
    e->p = take(Mm);
    e->r = take(Mm);
    e->Gp = take(Mm);
    take(Mm);
    e->dE = take(Nn);
    e->dc = take(Nn);
    e->tmp = take(Nn);
    e->partials = take((size_t)2 * kMaxRed * gmax + 64);
    e->dsc = take(ABIPGPU_SC_COUNT);
    e->dphase = take(32 + 2 * (size_t)gmax * kWarps + gmax);  // (+ the SM id of every CTA, debug builds)
    e->hsc = pinned_acquire();
    if (!e->hsc) return -1;

    LpCtx& c = e->ctx;
    c.m = (int)m;
    c.n = (int)n;
    c.A = Csr{e->A_ptr, e->A_idx, e->A_val, (int)m, e->A_wc, e->A_chunk, planA.lanes_log2,
              e->A_cl, e->A_lr, e->A_lp, 1, nullptr, nullptr, nullptr};
    c.AT = Csr{e->AT_ptr, e->AT_idx, e->AT_val, (int)n, e->AT_wc, e->AT_chunk, planAT.lanes_log2,
               e->AT_cl, e->AT_lr, e->AT_lp, 2, nullptr, nullptr, nullptr};
    c.M = e->dM;
    c.D = nullptr;
    c.E = nullptr;
    c.b = e->db;
    c.c = e->dc;
    c.h = e->vec[ABIPGPU_VEC_H];
    c.g = e->vec[ABIPGPU_VEC_G];
    c.rho_y = stgs->rho_y;
    c.alpha = stgs->alpha;
    c.cg_rate = stgs->cg_rate;
    c.g_th = 0.0;
    c.p = e->p;
    c.r = e->r;
    c.Gp = e->Gp;
    c.tmp = e->tmp;
    c.partials = e->partials;
    c.sc = e->dsc;
    memset(&c.comm, 0, sizeof(c.comm));
    c.comm.G = 1;

#ifdef ABIP_PHASE_TIMING
    c.phase_ns = e->dphase;
#else
    c.phase_ns = nullptr;
#endif

    lap(4);
    if (scale_out) {
        // equilibrate in the CALLER's order (bit-identical D, E to abip_normalize_A) on temporary copies of the original
        // structure, then gather the scaled values into the engine's two matrices
        int *d_perm = nullptr, *d_atptr = nullptr, *d_atidx = nullptr, *d_aptr = nullptr, *d_asrc = nullptr, *d_atsrc = nullptr;
        double* d_val = nullptr;
        std::vector<double> ax_copy(Ax, Ax + nnz);
        if (upload(&d_perm, perm, e) || upload(&d_atptr, at_ptr, e) || upload(&d_atidx, at_idx, e) || upload(&d_aptr, a_ptr, e) ||
            upload(&d_val, ax_copy, e))
            return -1;
        if (reorder && (upload(&d_asrc, e_a_src, e) || upload(&d_atsrc, e_at_src, e))) return -1;
        const unsigned gz = (unsigned)((nnz + 255) / 256);
        double* Dt = e->p;                    // scratch [m]
        double* rn = e->vec[ABIPGPU_VEC_UT];  // scratch [l] >= max(m, n)
        if (run_equilibration(e->stream, (int)m, (int)n, nnz, d_atptr, d_atidx, d_aptr, d_perm, d_val, e->dD, e->dE, Dt, rn, stgs,
                              scale_out->mean_row, scale_out->mean_col))
            return -1;
        k_gather_perm<<<gz, 256, 0, e->stream>>>(d_val, reorder ? d_asrc : d_perm, e->A_val, nnz);
        if (reorder) k_gather_perm<<<gz, 256, 0, e->stream>>>(d_val, d_atsrc, e->AT_val, nnz);
        else CK(cudaMemcpyAsync(e->AT_val, d_val, sizeof(double) * nnz, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemsetAsync(rn, 0, sizeof(double) * e->l, e->stream));
        CK(cudaMemsetAsync(Dt, 0, sizeof(double) * m, e->stream));
        CK(cudaMemcpyAsync(scale_out->D, e->dD, sizeof(double) * m, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaMemcpyAsync(scale_out->E, e->dE, sizeof(double) * n, cudaMemcpyDeviceToHost, e->stream));
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(e->stream));
        dev_free(d_perm, e->stream); dev_free(d_atptr, e->stream); dev_free(d_atidx, e->stream); dev_free(d_aptr, e->stream);
        dev_free(d_val, e->stream); dev_free(d_asrc, e->stream); dev_free(d_atsrc, e->stream);
        e->stats.d2h_bytes += 8.0 * 2 * (m + n);
    }
    lap(5);
    k_precond<<<(unsigned)((m + 255) / 256), 256, 0, e->stream>>>(c.A, e->dM);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(e->stream));
    lap(6);
    if (tune && tune_balance(e, e_a_ptr, e_at_ptr, &planA, &planAT, W, host_threads > 1)) return -1;
    lap(7);

    const double V = 8, I = 4;
    e->B_A = (double)nnz * (V + I) + (m + 1.0) * I + n * V + m * V;
    e->B_AT = (double)nnz * (V + I) + (n + 1.0) * I + m * V + n * V;
    snprintf(e->desc, sizeof(e->desc),
             "device %d (%s, %d SMs) persistent grid %d x %d threads, %zu B smem/block | CSR(A): %d rows, mean %.1f max %d "
             "nnz/row, %zu chunks (%d long rows), %d lane(s)/row | CSR(A'): %d rows, mean %.1f max %d, %zu chunks (%d long), "
             "%d lane(s)/row | nnz=%ld | locality ordering %s | set-up ms: transpose %.0f, ordering %.0f, permuted CSR %.0f, grid+plans %.0f, arena+upload %.0f, scaling %.0f, precond %.0f, measured balance %.0f (%d rounds, kept round %d: slowest CTA A' %.1f -> %.1f us, A %.1f -> %.1f us)",
             device, prop.name, e->num_sms, e->grid, kBlock, (size_t)e->smem, (int)m, planA.mean, planA.max_len,
             planA.chunk.size(), planA.n_long, 1 << planA.lanes_log2, (int)n, planAT.mean, planAT.max_len,
             planAT.chunk.size(), planAT.n_long, 1 << planAT.lanes_log2, nnz, reorder ? "on" : "off", e->setup_ms[0], e->setup_ms[1], e->setup_ms[2], e->setup_ms[3], e->setup_ms[4], e->setup_ms[5], e->setup_ms[6], e->setup_ms[7], e->tune_rounds, e->tune_best_round, e->tune_first_us[0], e->tune_best_us[0], e->tune_first_us[1], e->tune_best_us[1]);
    return 0;
}

static double account_solve(abipgpu_lp* e, double cg_its, bool warm) {
    // (c + 2 [+1 warm-start residual]) operator halves; 12 m-vector passes per CG iteration
    e->stats.n_solves++;
    e->stats.n_cg_iters += (abip_int)cg_its;
    const double nA = cg_its + 1 + (warm ? 1 : 0), nAT = cg_its + 1 + (warm ? 1 : 0);
    e->stats.n_spmv_A += (abip_int)nA;
    e->stats.n_spmv_AT += (abip_int)nAT;
    const double bytes = nA * e->B_A + nAT * e->B_AT + 12.0 * cg_its * e->m * 8.0;
    e->stats.alg_bytes += bytes;
    return bytes;
}

extern "C" {

abipgpu_lp* abipgpu_lp_create(abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai, const abip_float* Ax,
                              const ABIPSettings* stgs, int device) {
    if (!Ap || !Ai || !Ax || !stgs) return nullptr;
    abipgpu_lp* e = new abipgpu_lp();
    if (create_impl(e, m, n, Ap, Ai, Ax, stgs, device) != 0) {
        abipgpu_lp_destroy(e);
        return nullptr;
    }
    return e;
}

// same as abipgpu_lp_create but A is UNSCALED: the equilibration (common.c:150-565) runs on the device; D [m], E [n] and
// the two mean norms are returned to the host (needed for normalize_b_c and un_normalize_sol)
abipgpu_lp* abipgpu_lp_create_scaling(abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai, const abip_float* Ax,
                                      const ABIPSettings* stgs, int device, abip_float* D, abip_float* E,
                                      abip_float* mean_norm_row_A, abip_float* mean_norm_col_A) {
    if (!Ap || !Ai || !Ax || !stgs || !D || !E) return nullptr;
    abipgpu_lp* e = new abipgpu_lp();
    ScaleOut so{D, E, mean_norm_row_A, mean_norm_col_A};
    if (create_impl(e, m, n, Ap, Ai, Ax, stgs, device, &so) != 0) {
        abipgpu_lp_destroy(e);
        return nullptr;
    }
    return e;
}

// Equilibration of a whole matrix on the device without building an engine (multi-GPU set-up: every rank scales the FULL
// matrix on its own GPU instead of on its host cores -- common.c:150-565 took ~25 s per rank at cfg4).  Ax: in A, out the
// scaled matrix; D [m], E [n] and the two mean norms as abip_normalize_A returns them (bit-identical).
int abipgpu_equilibrate(abip_int m, abip_int n, const abip_int* Ap, const abip_int* Ai, abip_float* Ax, const ABIPSettings* stgs,
                        int device, abip_float* D, abip_float* E, abip_float* mean_norm_row_A, abip_float* mean_norm_col_A) {
    const long nnz = Ap[n];
    if (m <= 0 || n <= 0 || nnz <= 0 || nnz >= 2147483647L) return -1;
    CK(cudaSetDevice(device));
    const int host_threads = std::max(1, std::min(env_int("ABIP_GPU_HOST_THREADS", 8), (int)std::thread::hardware_concurrency()));
    std::vector<int> at_ptr, at_idx, a_ptr, a_idx, perm;
    if (transpose_csc(m, n, Ap, Ai, host_threads, &at_ptr, &at_idx, &a_ptr, &a_idx, &perm)) return -1;
    cudaStream_t st;
    CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    int *d_perm = nullptr, *d_atptr = nullptr, *d_atidx = nullptr, *d_aptr = nullptr;
    double *d_val = nullptr, *dD = nullptr, *dE = nullptr, *Dt = nullptr, *rn = nullptr;
    auto up = [&](auto** dst, const void* src, size_t bytes) -> int {
        CK(cudaMallocAsync((void**)dst, bytes + 64, st));
        if (src) CK(cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    };
    int rc = 0;
    if (up(&d_perm, perm.data(), sizeof(int) * nnz) || up(&d_atptr, at_ptr.data(), sizeof(int) * (n + 1)) ||
        up(&d_atidx, at_idx.data(), sizeof(int) * nnz) || up(&d_aptr, a_ptr.data(), sizeof(int) * (m + 1)) ||
        up(&d_val, Ax, sizeof(double) * nnz) || up(&dD, nullptr, sizeof(double) * m) || up(&dE, nullptr, sizeof(double) * n) ||
        up(&Dt, nullptr, sizeof(double) * m) || up(&rn, nullptr, sizeof(double) * std::max(m, n)))
        rc = -1;
    if (!rc) rc = run_equilibration(st, (int)m, (int)n, nnz, d_atptr, d_atidx, d_aptr, d_perm, d_val, dD, dE, Dt, rn, stgs,
                                   mean_norm_row_A, mean_norm_col_A);
    if (!rc) {
        if (cudaMemcpyAsync(Ax, d_val, sizeof(double) * nnz, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(D, dD, sizeof(double) * m, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaMemcpyAsync(E, dE, sizeof(double) * n, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
            cudaStreamSynchronize(st) != cudaSuccess)
            rc = -1;
    }
    void* ptrs[] = {d_perm, d_atptr, d_atidx, d_aptr, d_val, dD, dE, Dt, rn};
    for (void* q : ptrs)
        if (q) cudaFreeAsync(q, st);
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    return rc;
}

void abipgpu_lp_destroy(abipgpu_lp* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    {
        void* ptrs[] = {e->arena, e->plan_arena, e->slab, e->d_pl, e->d_rn2o, e->d_cn2o};
        for (void* q : ptrs) {
            if (e->stream && q) dev_free(q, e->stream);  // stream-ordered: no device-wide synchronisation
        }
    }
    for (int q = 0; q < kMaxRanks; ++q)
        if (e->peer_bufs[q] && q != e->dist_rank) cudaIpcCloseMemHandle(e->peer_bufs[q]);
    if (e->comm_buf) cudaFree(e->comm_buf);
    if (e->hsc) pinned_release(e->hsc);
    if (e->ev0) cudaEventDestroy(e->ev0);
    if (e->ev1) cudaEventDestroy(e->ev1);
    if (e->ev_solve0) cudaEventDestroy(e->ev_solve0);
    if (e->ev_solve1) cudaEventDestroy(e->ev_solve1);
    if (e->stream && e->own_stream) cudaStreamDestroy(e->stream);
    delete e;
}

// g = K^-1 h, g_x *= -1, g_th = h.g (src/abip.c:1915-1924) as a launch of its own
static int solve_g_now(abipgpu_lp* e) {
    if (launch_coop(e, e->dist_G > 1 ? (const void*)k_solve_vec<true> : (const void*)k_solve_vec<false>, e->grid, e->smem, e->ctx, e->vec[ABIPGPU_VEC_G], (const double*)nullptr,
                    (long)-1, 1))
        return -1;
    if (read_sc(e, nullptr)) return -1;
    e->ctx.g_th = e->hsc[ABIPGPU_SC_VEC_NORM2];
    e->need_g = false;
    account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS], false);
    return 0;
}

int abipgpu_lp_set_problem(abipgpu_lp* e, const abip_float* b, const abip_float* c, const abip_float* D,
                           const abip_float* E) {
    CK(cudaSetDevice(e->device));
    const int m = e->m, n = e->n;
    e->have_scaling = (D && E);
    if (!e->permuted) {
        CK(cudaMemcpyAsync(e->db, b, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->dc, c, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
        if (e->have_scaling) {
            CK(cudaMemcpyAsync(e->dD, D, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
            CK(cudaMemcpyAsync(e->dE, E, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
        }
    } else {  // caller's order -> engine order
        const unsigned gm = (unsigned)((m + 255) / 256), gn = (unsigned)((n + 255) / 256);
        CK(cudaMemcpyAsync(e->xin, b, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
        CK(cudaMemcpyAsync(e->xin + m, c, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
        k_perm_in<<<gm, 256, 0, e->stream>>>(e->xin, e->d_rn2o, e->db, m, m);
        k_perm_in<<<gn, 256, 0, e->stream>>>(e->xin + m, e->d_cn2o, e->dc, n, n);
        if (e->have_scaling) {
            CK(cudaMemcpyAsync(e->yout, D, sizeof(double) * m, cudaMemcpyHostToDevice, e->stream));
            CK(cudaMemcpyAsync(e->yout + m, E, sizeof(double) * n, cudaMemcpyHostToDevice, e->stream));
            k_perm_in<<<gm, 256, 0, e->stream>>>(e->yout, e->d_rn2o, e->dD, m, m);
            k_perm_in<<<gn, 256, 0, e->stream>>>(e->yout + m, e->d_cn2o, e->dE, n, n);
        }
        CK(cudaGetLastError());
    }
    e->stats.h2d_bytes += 8.0 * (m + n) * (e->have_scaling ? 2 : 1);
    e->ctx.D = e->have_scaling ? e->dD : nullptr;
    e->ctx.E = e->have_scaling ? e->dE : nullptr;
    k_build_h<<<(m + n + 255) / 256, 256, 0, e->stream>>>(e->db, e->dc, e->vec[ABIPGPU_VEC_H], e->vec[ABIPGPU_VEC_G], m, n);
    CK(cudaGetLastError());
    if (e->batch) {  // batch engines: inside the first batched step of the problem (k_batch, BatchItem::solve_g)
        e->dirty = true;
        e->need_g = true;
        e->ctx.g_th = 0.0;
        return 0;
    }
    return solve_g_now(e);
}

abip_float abipgpu_lp_g_th(const abipgpu_lp* e) { return e->ctx.g_th; }

int abipgpu_lp_cold_start(abipgpu_lp* e, abip_float mu, abip_float beta) {
    e->dirty = true;
    if (e->batch) return defer(e, PreOp{PRE_COLD, 0, 0, 0, mu, beta});
    CK(cudaSetDevice(e->device));
    k_cold_start<<<(e->l + 255) / 256, 256, 0, e->stream>>>(e->vec[ABIPGPU_VEC_U], e->vec[ABIPGPU_VEC_V], e->m, e->l,
                                                             sqrt(mu / beta));
    CK(cudaGetLastError());
    return 0;
}

int abipgpu_lp_outer_prologue(abipgpu_lp* e, int avg_criterion) {
    e->dirty = true;
    if (e->batch) {
        e->restart_synced = true;
        return defer(e, PreOp{PRE_PROLOGUE, avg_criterion, 0, 0, 0.0, 0.0});
    }
    CK(cudaSetDevice(e->device));
    const size_t bytes = sizeof(double) * e->l;
    CK(cudaMemsetAsync(e->vec[ABIPGPU_VEC_USUM], 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->vec[ABIPGPU_VEC_VSUM], 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->vec[ABIPGPU_VEC_UAVG], 0, bytes, e->stream));
    CK(cudaMemsetAsync(e->vec[ABIPGPU_VEC_VAVG], 0, bytes, e->stream));
    e->restart_synced = true;  // u_avg == u_sum == 0 at the start of an outer iteration
    if (avg_criterion) {
        CK(cudaMemcpyAsync(e->vec[ABIPGPU_VEC_U], e->vec[ABIPGPU_VEC_UAVGC], bytes, cudaMemcpyDeviceToDevice, e->stream));
        CK(cudaMemcpyAsync(e->vec[ABIPGPU_VEC_V], e->vec[ABIPGPU_VEC_VAVGC], bytes, cudaMemcpyDeviceToDevice, e->stream));
    }
    return 0;
}

int abipgpu_lp_admm_iter(abipgpu_lp* e, abip_int j, abip_int k, abip_float mu, abip_float beta, abip_float* sc) {
    CK(cudaSetDevice(e->device));
    IterArgs a;
    a.u = e->vec[ABIPGPU_VEC_U];
    a.v = e->vec[ABIPGPU_VEC_V];
    a.ut = e->vec[ABIPGPU_VEC_UT];
    a.u_prev = e->vec[ABIPGPU_VEC_UPREV];
    a.u_sum = e->vec[ABIPGPU_VEC_USUM];
    a.v_sum = e->vec[ABIPGPU_VEC_VSUM];
    a.u_avgc = e->vec[ABIPGPU_VEC_UAVGC];
    a.v_avgc = e->vec[ABIPGPU_VEC_VAVGC];
    a.u_avg = e->vec[ABIPGPU_VEC_UAVG];
    a.v_avg = e->vec[ABIPGPU_VEC_VAVG];
    a.j = j;
    a.k = k;
    a.mu = mu;
    a.beta = beta;
    a.half_update = (int)e->stgs.half_update;
    // restart_vars (abip.c:587-630): u_avg/v_avg equal u_sumcon/v_sumcon until the first restart of an outer
    // iteration fires, and are only read once k >= restart_thresh, so they are materialised lazily.
    a.restart_active = (k >= e->stgs.restart_thresh) ? 1 : 0;
    a.restart_fire = 0;
    a.restart_fre = (double)e->stgs.restart_fre;
    if (a.restart_active) {
        if (e->restart_synced && j > 0) {
            const size_t bytes = sizeof(double) * e->l;
            if (e->batch) {
                if (defer(e, PreOp{PRE_RESTART_COPY, 0, 0, 0, 0.0, 0.0})) return -1;
            } else {
                CK(cudaMemcpyAsync(a.u_avg, a.u_sum, bytes, cudaMemcpyDeviceToDevice, e->stream));
                CK(cudaMemcpyAsync(a.v_avg, a.v_sum, bytes, cudaMemcpyDeviceToDevice, e->stream));
            }
        }
        e->restart_synced = false;
    }
    // fre_old is 0 until the first restart of this outer iteration and fre afterwards (abip.c:628, 2116):
    // both give the same residue class, so the firing rule reduces to (j+1) % fre == 0
    if (a.restart_active && e->stgs.restart_fre > 0 && (j + 1) % e->stgs.restart_fre == 0) a.restart_fire = 1;
    if (e->batch) {  // lock-step batch: the executor launches this step together with those of the other problems
        BatchReq r;
        r.e = e;
        r.item.c = e->ctx;
        r.item.kind = BATCH_ADMM;
        r.item.it = a;
        if (batch_step(e, &r, sc)) return -1;
    } else {
        CK(cudaEventRecord(e->ev0, e->stream));
        if (launch_coop(e, e->dist_G > 1 ? (const void*)k_admm_iter<true> : (const void*)k_admm_iter<false>, e->grid, e->smem, e->ctx, a)) return -1;
        CK(cudaEventRecord(e->ev1, e->stream));
        if (read_sc(e, sc)) return -1;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        e->stats.admm_kernel_ms += ms;
    }
    e->stats.n_admm_launch++;
    double bytes = account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS], true);
    const double nq = e->hsc[ABIPGPU_SC_HAS_AVG] != 0 ? 2.0 : 1.0;
    e->stats.n_spmv_A += (abip_int)nq;
    e->stats.n_spmv_AT += (abip_int)nq;
    // rhs (6 l) + prox/dual/avg (9 l) + Q-norm extra vectors (b, c, D, E, s, y: ~3 l) per pair
    const double extra = nq * (e->B_A + e->B_AT) + (15.0 + 3.0 * nq) * e->l * 8.0;
    e->stats.alg_bytes += extra;
    e->stats.alg_bytes_admm += bytes + extra;
    return 0;
}

int abipgpu_lp_mu_stats(abipgpu_lp* e, int avg_criterion, abip_float* sc) {
    CK(cudaSetDevice(e->device));
    const double* u = e->vec[avg_criterion ? ABIPGPU_VEC_UAVGC : ABIPGPU_VEC_U];
    const double* v = e->vec[avg_criterion ? ABIPGPU_VEC_VAVGC : ABIPGPU_VEC_V];
    if (e->batch) {
        BatchReq r;
        r.e = e;
        r.item.c = e->ctx;
        r.item.kind = BATCH_MU;
        r.item.mu.u = u;
        r.item.mu.v = v;
        return batch_step(e, &r, sc);
    }
    if (launch_coop(e, (const void*)k_mu_stats, e->grid_mu, (size_t)0, u, v, e->m, e->l, e->partials, e->dsc, e->ctx.comm)) return -1;
    return read_sc(e, sc);
}

int abipgpu_lp_reinit(abipgpu_lp* e, int indx, abip_float sigma, int avg_criterion) {
    e->dirty = true;
    if (e->batch) return defer(e, PreOp{PRE_REINIT, indx, avg_criterion, 0, sigma, 0.0});
    CK(cudaSetDevice(e->device));
    double* u = e->vec[avg_criterion ? ABIPGPU_VEC_UAVGC : ABIPGPU_VEC_U];
    double* v = e->vec[avg_criterion ? ABIPGPU_VEC_VAVGC : ABIPGPU_VEC_V];
    k_reinit<<<(e->n + 1 + 255) / 256, 256, 0, e->stream>>>(u, v, e->m, e->l, indx, sigma);
    CK(cudaGetLastError());
    e->stats.n_kernel_launches++;
    return 0;
}

int abipgpu_lp_clamp_v(abipgpu_lp* e) {
    e->dirty = true;
    if (e->batch) return defer(e, PreOp{PRE_CLAMP, 0, 0, 0, 0.0, 0.0});
    CK(cudaSetDevice(e->device));
    k_clamp_v<<<(e->l + 255) / 256, 256, 0, e->stream>>>(e->vec[ABIPGPU_VEC_V], e->l);
    CK(cudaGetLastError());
    e->stats.n_kernel_launches++;
    return 0;
}

int abipgpu_lp_bb_begin(abipgpu_lp* e) {
    e->dirty = true;
    if (e->batch) return defer(e, PreOp{PRE_BB_BEGIN, 0, 0, 0, 0.0, 0.0});
    CK(cudaSetDevice(e->device));
    const size_t bytes = sizeof(double) * e->l;
    CK(cudaMemcpyAsync(e->vec[ABIPGPU_VEC_BB_UPREV], e->vec[ABIPGPU_VEC_U], bytes, cudaMemcpyDeviceToDevice, e->stream));
    CK(cudaMemcpyAsync(e->vec[ABIPGPU_VEC_BB_VPREV], e->vec[ABIPGPU_VEC_V], bytes, cudaMemcpyDeviceToDevice, e->stream));
    return 0;
}

int abipgpu_lp_bb_round(abipgpu_lp* e, int carry, abip_int k, abip_float mu, abip_float beta_prev, abip_float* sc) {
    CK(cudaSetDevice(e->device));
    BBArgs a;
    a.u_prev = e->vec[ABIPGPU_VEC_BB_UPREV];
    a.v_prev = e->vec[ABIPGPU_VEC_BB_VPREV];
    a.ut = e->vec[ABIPGPU_VEC_BB_UT];
    a.u = e->vec[ABIPGPU_VEC_BB_U];
    a.v = e->vec[ABIPGPU_VEC_BB_V];
    a.ut_next = e->vec[ABIPGPU_VEC_BB_UTNEXT];
    a.u_next = e->vec[ABIPGPU_VEC_BB_UNEXT];
    a.v_next = e->vec[ABIPGPU_VEC_BB_VNEXT];
    a.carry = carry;
    a.k = k;
    a.mu = mu;
    a.beta_prev = beta_prev;
    if (e->batch) {
        BatchReq r;
        r.e = e;
        r.item.c = e->ctx;
        r.item.kind = BATCH_BB;
        r.item.bb = a;
        if (batch_step(e, &r, sc)) return -1;
    } else {
        CK(cudaEventRecord(e->ev0, e->stream));
        if (launch_coop(e, e->dist_G > 1 ? (const void*)k_bb_round<true> : (const void*)k_bb_round<false>, e->grid, e->smem, e->ctx, a)) return -1;
        CK(cudaEventRecord(e->ev1, e->stream));
        if (read_sc(e, sc)) return -1;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
        e->stats.bb_kernel_ms += ms;
    }
    e->stats.n_bb_launch++;
    double bytes = account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS], true);
    bytes += account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS2], true);
    const double extra = (2 * 6.0 + 2 * 6.0 + (carry ? 4.0 : 0.0)) * e->l * 8.0;
    e->stats.alg_bytes += extra;
    e->stats.alg_bytes_bb += bytes + extra;
    return 0;
}

// ---- device-resident loops (batch engines only): one batched step = the inner ADMM loop of an outer iteration (up to
// L->cap iterations) or a whole Barzilai-Borwein search
int abipgpu_lp_is_batch(const abipgpu_lp* e) { return e->batch != nullptr; }

static void fill_iter_args(abipgpu_lp* e, IterArgs* a) {
    a->u = e->vec[ABIPGPU_VEC_U]; a->v = e->vec[ABIPGPU_VEC_V]; a->ut = e->vec[ABIPGPU_VEC_UT];
    a->u_prev = e->vec[ABIPGPU_VEC_UPREV]; a->u_sum = e->vec[ABIPGPU_VEC_USUM]; a->v_sum = e->vec[ABIPGPU_VEC_VSUM];
    a->u_avgc = e->vec[ABIPGPU_VEC_UAVGC]; a->v_avgc = e->vec[ABIPGPU_VEC_VAVGC];
    a->u_avg = e->vec[ABIPGPU_VEC_UAVG]; a->v_avg = e->vec[ABIPGPU_VEC_VAVG];
    a->half_update = (int)e->stgs.half_update;
    a->restart_active = 0; a->restart_fire = 0; a->restart_fre = (double)e->stgs.restart_fre;
}

int abipgpu_lp_inner_loop(abipgpu_lp* e, const LpInnerArgs* L, abip_float* sc) {
    if (!e->batch) return -1;
    BatchReq r;
    r.e = e;
    r.item.c = e->ctx;
    r.item.kind = BATCH_INNER;
    fill_iter_args(e, &r.item.it);
    r.item.it.j = L->j0; r.item.it.k = L->k0; r.item.it.mu = L->mu; r.item.it.beta = L->beta;
    memset(&r.item.solve, 0, sizeof(r.item.solve));
    r.item.solve.in = *L;
    if (batch_step(e, &r, sc)) return -1;
    e->stats.n_admm_launch++;
    return 0;
}

static void fill_bb_args(abipgpu_lp* e, BBArgs* a) {
    a->u_prev = e->vec[ABIPGPU_VEC_BB_UPREV]; a->v_prev = e->vec[ABIPGPU_VEC_BB_VPREV]; a->ut = e->vec[ABIPGPU_VEC_BB_UT];
    a->u = e->vec[ABIPGPU_VEC_BB_U]; a->v = e->vec[ABIPGPU_VEC_BB_V]; a->ut_next = e->vec[ABIPGPU_VEC_BB_UTNEXT];
    a->u_next = e->vec[ABIPGPU_VEC_BB_UNEXT]; a->v_next = e->vec[ABIPGPU_VEC_BB_VNEXT];
    a->carry = 0; a->k = 0; a->mu = 1.0; a->beta_prev = 1.0;
}

// the device-resident OUTER loop (LpSolveArgs, lp_logic.h): one batched step runs the solve until it ends, the launch cap is
// reached or the host is needed; the state comes back in sc[ABIPGPU_SC_LOOP_*]
int abipgpu_lp_solve_loop(abipgpu_lp* e, const LpSolveArgs* S, abip_float* sc) {
    if (!e->batch) return -1;
    BatchReq r;
    r.e = e;
    r.item.c = e->ctx;
    r.item.kind = BATCH_SOLVE;
    fill_iter_args(e, &r.item.it);
    r.item.it.j = S->in.j0; r.item.it.k = S->in.k0; r.item.it.mu = S->in.mu; r.item.it.beta = S->in.beta;
    fill_bb_args(e, &r.item.bb);
    r.item.solve = *S;
    r.item.search = BBSearchArgs{S->adaptive_lookback, S->eps_cor, S->eps_pen};
    if (batch_step(e, &r, sc)) return -1;
    // the device ran the prologue of the current outer iteration (u_sum = u_avg = 0) and leaves before the restart
    // bookkeeping starts: the lazily materialised u_avg / v_avg of abipgpu_lp_admm_iter are still in sync
    e->restart_synced = true;
    e->stats.n_admm_launch++;
    return 0;
}

int abipgpu_lp_bb_search(abipgpu_lp* e, abip_int k, abip_float mu, int lookback, abip_float eps_cor, abip_float eps_pen,
                         abip_float* sc) {
    if (!e->batch) return -1;
    BatchReq r;
    r.e = e;
    r.item.c = e->ctx;
    r.item.kind = BATCH_BBSEARCH;
    BBArgs& a = r.item.bb;
    fill_bb_args(e, &a);
    a.k = k; a.mu = mu;
    memset(&r.item.solve, 0, sizeof(r.item.solve));
    r.item.search = BBSearchArgs{lookback, eps_cor, eps_pen};
    if (batch_step(e, &r, sc)) return -1;
    e->stats.n_bb_launch++;
    return 0;
}

int abipgpu_lp_solve_vec(abipgpu_lp* e, int rhs_id, int warm_id, abip_int iter, abip_float* sc) {
    if (flush_pending(e)) return -1;
    CK(cudaSetDevice(e->device));
    if (rhs_id < 0 || rhs_id > 20 || warm_id > 20) return -1;
    const double* s = warm_id >= 0 ? e->vec[warm_id] : nullptr;
    if (launch_coop(e, e->dist_G > 1 ? (const void*)k_solve_vec<true> : (const void*)k_solve_vec<false>, e->grid, e->smem, e->ctx, e->vec[rhs_id], s, (long)iter, 0)) return -1;
    if (read_sc(e, sc)) return -1;
    account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS], s != nullptr);
    return 0;
}

}  // extern "C"
// executes the deferred operations of a batch engine with ordinary launches (needed before anything other than a
// batched step reads the vectors: get/set_vec, solve_vec)
static int flush_pending(abipgpu_lp* e) {
    if (e->need_g) {  // something other than a batched step is about to read the vectors: compute g with a launch of its own
        if (e->dirty) {
            CK(cudaStreamSynchronize(e->stream));
            e->dirty = false;
        }
        if (solve_g_now(e)) return -1;
    }
    if (!e->n_pend) return 0;
    BatchExec* b = e->batch;
    const bool synced = e->restart_synced;
    e->batch = nullptr;
    int rc = 0;
    const int n = e->n_pend;
    e->n_pend = 0;
    for (int q = 0; q < n && !rc; ++q) {
        const PreOp& op = e->pend[q];
        if (op.code == PRE_COLD) rc = abipgpu_lp_cold_start(e, op.d0, op.d1);
        else if (op.code == PRE_PROLOGUE) rc = abipgpu_lp_outer_prologue(e, op.i0);
        else if (op.code == PRE_REINIT) rc = abipgpu_lp_reinit(e, op.i0, op.d0, op.i1);
        else if (op.code == PRE_CLAMP) rc = abipgpu_lp_clamp_v(e);
        else if (op.code == PRE_BB_BEGIN) rc = abipgpu_lp_bb_begin(e);
        else if (op.code == PRE_RESTART_COPY) {
            const size_t bytes = sizeof(double) * e->l;
            if (cudaMemcpyAsync(e->vec[ABIPGPU_VEC_UAVG], e->vec[ABIPGPU_VEC_USUM], bytes, cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess ||
                cudaMemcpyAsync(e->vec[ABIPGPU_VEC_VAVG], e->vec[ABIPGPU_VEC_VSUM], bytes, cudaMemcpyDeviceToDevice, e->stream) != cudaSuccess)
                rc = -1;
        }
    }
    e->restart_synced = synced;
    e->batch = b;
    e->dirty = true;
    return rc;
}
extern "C" {
int abipgpu_lp_get_vec(abipgpu_lp* e, int id, abip_float* host, abip_int len) {
    if (flush_pending(e)) return -1;
    CK(cudaSetDevice(e->device));
    if (id < 0 || id > 20 || len > e->vec_len[id]) return -1;
    if (e->permuted) {  // engine order -> caller's order
        const int vl = (int)e->vec_len[id];
        k_perm_out<<<(unsigned)((vl + 255) / 256), 256, 0, e->stream>>>(e->vec[id], id == ABIPGPU_VEC_M ? e->d_rn2o : e->d_pl, e->yout, vl);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(host, e->yout, sizeof(double) * len, cudaMemcpyDeviceToHost, e->stream));
    } else {
        CK(cudaMemcpyAsync(host, e->vec[id], sizeof(double) * len, cudaMemcpyDeviceToHost, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    e->stats.d2h_bytes += 8.0 * len;
    return 0;
}

int abipgpu_lp_set_vec(abipgpu_lp* e, int id, const abip_float* host, abip_int len) {
    if (flush_pending(e)) return -1;
    CK(cudaSetDevice(e->device));
    if (id < 0 || id > 20 || len > e->vec_len[id]) return -1;
    if (e->permuted) {  // caller's order -> engine order (entries beyond len keep their values)
        const int vl = (int)e->vec_len[id];
        CK(cudaMemcpyAsync(e->xin, host, sizeof(double) * len, cudaMemcpyHostToDevice, e->stream));
        k_perm_in<<<(unsigned)((vl + 255) / 256), 256, 0, e->stream>>>(e->xin, id == ABIPGPU_VEC_M ? e->d_rn2o : e->d_pl, e->vec[id], vl, (int)len);
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(e->vec[id], host, sizeof(double) * len, cudaMemcpyHostToDevice, e->stream));
    }
    CK(cudaStreamSynchronize(e->stream));
    e->stats.h2d_bytes += 8.0 * len;
    return 0;
}

int abipgpu_lp_spmv(abipgpu_lp* e, int trans, const abip_float* x, abip_float* y) {
    return abipgpu_lp_spmv_host(e, trans, x, y, 0);
}

abip_int abipgpu_plan_debug(abip_int nrows, const int* rowptr, abip_int ctas, abip_int deal, int* chunks4,
                            abip_int max_chunks, int* warp_chunk, int* info6) {
    return abipgpu_plan_debug_cost(nrows, rowptr, ctas, deal, nullptr, chunks4, max_chunks, warp_chunk, info6);
}

// same with explicit per-row costs for the cut of the CTA row ranges (the path of the measured balance, tune_balance)
abip_int abipgpu_plan_debug_cost(abip_int nrows, const int* rowptr, abip_int ctas, abip_int deal, const double* row_cost,
                                 int* chunks4, abip_int max_chunks, int* warp_chunk, int* info6) {
    if (nrows <= 0 || !rowptr || ctas <= 0) return 0;
    std::vector<int> ptr(rowptr, rowptr + nrows + 1);
    std::vector<double> cost;
    if (row_cost) cost.assign(row_cost, row_cost + nrows);
    SpmvPlan P;
    const char* prev = getenv("ABIP_GPU_PLAN_DEAL");
    const std::string saved = prev ? prev : "";
    setenv("ABIP_GPU_PLAN_DEAL", deal ? "1" : "0", 1);
    build_spmv_plan(ptr, (int)nrows, (int)ctas * kWarps, "ABIP_GPU_LANES_DEBUG", &P, nullptr, row_cost ? &cost : nullptr);
    if (prev) setenv("ABIP_GPU_PLAN_DEAL", saved.c_str(), 1);
    else unsetenv("ABIP_GPU_PLAN_DEAL");
    if (info6) {
        info6[0] = kWarps; info6[1] = kChunk; info6[2] = kChunkRows;
        info6[3] = P.n_long; info6[4] = P.n_pieces; info6[5] = P.lanes_log2;
    }
    if ((abip_int)P.chunk.size() > max_chunks) return -(abip_int)P.chunk.size();
    for (size_t i = 0; i < P.chunk.size(); ++i) {
        chunks4[4 * i] = P.chunk[i].x; chunks4[4 * i + 1] = P.chunk[i].y;
        chunks4[4 * i + 2] = P.chunk[i].z; chunks4[4 * i + 3] = P.chunk[i].w;
    }
    if (warp_chunk) memcpy(warp_chunk, P.warp_chunk.data(), sizeof(int) * P.warp_chunk.size());
    return (abip_int)P.chunk.size();
}

void abipgpu_lp_describe(const abipgpu_lp* e, char* buf, abip_int buflen) {
    if (buflen > 0) snprintf(buf, (size_t)buflen, "%s", e->desc);
}

}  // extern "C"

// ---- internal helpers shared with lp_host.cpp -------------------------------------------------------------
int abipgpu_lp_spmv_host(abipgpu_lp* e, int trans, const double* x, double* y, int accumulate) {
    CK(cudaSetDevice(e->device));
    const int nin = trans ? e->m : e->n, nout = trans ? e->n : e->m;
    CK(cudaMemcpyAsync(e->xin, x, sizeof(double) * nin, cudaMemcpyHostToDevice, e->stream));
    if (accumulate) CK(cudaMemcpyAsync(e->yout, y, sizeof(double) * nout, cudaMemcpyHostToDevice, e->stream));
    if (e->permuted) {
        const int* pin = trans ? e->d_rn2o : e->d_cn2o;
        const int* pout = trans ? e->d_cn2o : e->d_rn2o;
        k_perm_in<<<(unsigned)((nin + 255) / 256), 256, 0, e->stream>>>(e->xin, pin, e->xin2, nin, nin);
        if (accumulate) k_perm_in<<<(unsigned)((nout + 255) / 256), 256, 0, e->stream>>>(e->yout, pout, e->yout2, nout, nout);
        k_spmv<<<e->grid, kBlock, e->smem, e->stream>>>(trans ? e->ctx.AT : e->ctx.A, e->xin2, e->yout2, accumulate);
        k_perm_out<<<(unsigned)((nout + 255) / 256), 256, 0, e->stream>>>(e->yout2, pout, e->yout, nout);
    } else {
        k_spmv<<<e->grid, kBlock, e->smem, e->stream>>>(trans ? e->ctx.AT : e->ctx.A, e->xin, e->yout, accumulate);
    }
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(y, e->yout, sizeof(double) * nout, cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    e->stats.n_kernel_launches++;
    e->stats.h2d_bytes += 8.0 * (nin + (accumulate ? nout : 0));
    e->stats.d2h_bytes += 8.0 * nout;
    if (trans) e->stats.n_spmv_AT++; else e->stats.n_spmv_A++;
    return 0;
}

// solve_lin_sys with host pointers (plugin path): b [m+n] in/out, s [m] or NULL
int abipgpu_lp_solve_host(abipgpu_lp* e, double* b, const double* s, long iter, int* cg_its) {
    CK(cudaSetDevice(e->device));
    const int mn = e->m + e->n;
    CK(cudaMemcpyAsync(e->xin, b, sizeof(double) * mn, cudaMemcpyHostToDevice, e->stream));
    if (s) CK(cudaMemcpyAsync(e->yout, s, sizeof(double) * e->m, cudaMemcpyHostToDevice, e->stream));
    double* rhs = e->xin;
    const double* warm = s ? e->yout : nullptr;
    if (e->permuted) {
        k_perm_in<<<(unsigned)((mn + 255) / 256), 256, 0, e->stream>>>(e->xin, e->d_pl, e->xin2, mn, mn);
        if (s) k_perm_in<<<(unsigned)((e->m + 255) / 256), 256, 0, e->stream>>>(e->yout, e->d_rn2o, e->yout2, e->m, e->m);
        CK(cudaGetLastError());
        rhs = e->xin2;
        warm = s ? e->yout2 : nullptr;
    }
    if (launch_coop(e, e->dist_G > 1 ? (const void*)k_solve_vec<true> : (const void*)k_solve_vec<false>, e->grid, e->smem, e->ctx, rhs, warm,
                    iter, 0))
        return -1;
    if (e->permuted) {
        k_perm_out<<<(unsigned)((mn + 255) / 256), 256, 0, e->stream>>>(e->xin2, e->d_pl, e->xin, mn);
        CK(cudaGetLastError());
    }
    CK(cudaMemcpyAsync(b, e->xin, sizeof(double) * mn, cudaMemcpyDeviceToHost, e->stream));
    if (read_sc(e, nullptr)) return -1;
    e->stats.h2d_bytes += 8.0 * (mn + (s ? e->m : 0));
    e->stats.d2h_bytes += 8.0 * mn;
    account_solve(e, e->hsc[ABIPGPU_SC_CG_ITS], s != nullptr);
    if (cg_its) *cg_its = (int)e->hsc[ABIPGPU_SC_CG_ITS];
    return 0;
}

// ---- multi-GPU communicator set-up (CUDA IPC): each rank allocates one buffer, exports its handle, and maps the
// buffers of all peers.  The exchange of the 64-byte handles is the caller's job (torch.distributed all_gather).
static void comm_layout(long m, size_t* off_scal, size_t* off_flags, size_t* off_seq, size_t* off_err, size_t* total,
                        long* m_pad) {
    *m_pad = (m + 31) & ~31L;
    *off_scal = sizeof(double) * 3 * (size_t)*m_pad;  // [2][m_pad] partials + [m_pad] reduced slice
    *off_flags = *off_scal + sizeof(double) * 2 * kCommScalars;
    *off_seq = *off_flags + sizeof(unsigned long long) * kMaxRanks;
    *off_err = *off_seq + sizeof(unsigned long long);
    *total = ((*off_err + sizeof(int) + 255) / 256) * 256;
}

extern "C" int abipgpu_lp_comm_export(abipgpu_lp* e, void* handle64) {
    CK(cudaSetDevice(e->device));
    size_t o1, o2, o3, o4, total;
    long m_pad;
    comm_layout(e->m, &o1, &o2, &o3, &o4, &total, &m_pad);
    if (!e->comm_buf) {
        CK(cudaMalloc((void**)&e->comm_buf, total));
        CK(cudaMemset(e->comm_buf, 0, total));
        e->comm_bytes = total;
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, e->comm_buf));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
    memcpy(handle64, &h, 64);
    return 0;
}

extern "C" int abipgpu_lp_comm_connect(abipgpu_lp* e, int G, int rank, const void* handles /* G x 64 bytes */) {
    CK(cudaSetDevice(e->device));
    if (G < 1 || G > kMaxRanks || rank < 0 || rank >= G || !e->comm_buf) return -1;
    size_t o1, o2, o3, o4, total;
    long m_pad;
    comm_layout(e->m, &o1, &o2, &o3, &o4, &total, &m_pad);
    Comm& cm = e->ctx.comm;
    cm.G = G;
    cm.rank = rank;
    cm.m_pad = m_pad;
    for (int q = 0; q < G; ++q) {
        void* base = nullptr;
        if (q == rank) {
            base = e->comm_buf;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, (const unsigned char*)handles + 64 * q, 64);
            CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
        }
        e->peer_bufs[q] = base;
        cm.vec[q] = (double*)base;
        cm.red[q] = (double*)base + 2 * m_pad;
        cm.scal[q] = (double*)((unsigned char*)base + o1);
        cm.flags[q] = (unsigned long long*)((unsigned char*)base + o2);
    }
    cm.seq = (unsigned long long*)(e->comm_buf + o3);
    cm.err = (int*)(e->comm_buf + o4);
    e->dist_G = G;
    e->dist_rank = rank;
    return 0;
}

extern "C" void abipgpu_lp_set_global_n(abipgpu_lp* e, long n_global) { e->n_global = n_global; }

ABIPGpuStats* abipgpu_lp_stats(abipgpu_lp* e) { return &e->stats; }
extern "C" int abipgpu_lp_phase_times(abipgpu_lp* e, double* out32, int reset) {  // debug builds (-DABIP_PHASE_TIMING)
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(out32, e->dphase, sizeof(double) * 32, cudaMemcpyDeviceToHost, e->stream));
    if (reset) CK(cudaMemsetAsync(e->dphase, 0, sizeof(double) * 32, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
extern "C" int abipgpu_lp_warp_times(abipgpu_lp* e, double* out, int* W) {  // [2][W] per-warp SpMV busy ns (debug builds)
    CK(cudaSetDevice(e->device));
    *W = e->grid * kWarps;
    CK(cudaMemcpyAsync(out, e->dphase + 32, sizeof(double) * (2 * (*W) + e->grid), cudaMemcpyDeviceToHost, e->stream));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
#ifdef ABIP_PHASE_TIMING
extern "C" int abipgpu_lp_spmv_prof(abipgpu_lp* e, unsigned long long* out16, int reset) {
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    CK(cudaMemcpyFromSymbol(out16, g_spmv_prof, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0};
        CK(cudaMemcpyToSymbol(g_spmv_prof, z, sizeof(z)));
    }
    return 0;
}
#endif
int abipgpu_lp_solve_timer(abipgpu_lp* e, int stop) {  // CUDA events on the engine stream around a whole solve
    if (e->batch) return 0;  // batched steps run on the executor's stream; the host clock times the batch
    CK(cudaSetDevice(e->device));
    if (!stop) {
        CK(cudaEventRecord(e->ev_solve0, e->stream));
    } else {
        CK(cudaEventRecord(e->ev_solve1, e->stream));
        CK(cudaEventSynchronize(e->ev_solve1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e->ev_solve0, e->ev_solve1));
        e->stats.solve_event_ms = ms;
    }
    return 0;
}
int abipgpu_lp_dims(const abipgpu_lp* e, int* m, int* n) { *m = e->m; *n = e->n; return 0; }
void abipgpu_lp_drop_pending(abipgpu_lp* e) { e->n_pend = 0; }
int abipgpu_lp_sync(abipgpu_lp* e) {
    CK(cudaSetDevice(e->device));
    CK(cudaStreamSynchronize(e->stream));
    return 0;
}
