// lp_device.cuh -- device-side building blocks of the ABIP-LP engine (sm_100a).
//
// Everything here is a __device__ "phase": a grid-wide loop executed by all threads of ONE persistent
// cooperative kernel (one launch = one ADMM iteration / one BB round / one linear solve), separated by
// grid barriers.  All kernels are HBM-bound FP64 streaming/gather work; tensor cores are not used.
//
// Determinism: every reduction is (thread-serial) -> (warp shuffle tree) -> (per-block shared memory, fixed
// order) -> per-block partial in global memory -> after the grid barrier EVERY block sums the G partials in the
// same fixed order, so all blocks hold bit-identical scalars and take identical branches (a requirement for
// the data-dependent PCG loop to be barrier-safe).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>

namespace cg = cooperative_groups;

#ifndef ABIP_BLOCK
#define ABIP_BLOCK 512
#endif
#ifndef ABIP_MIN_BLOCKS_PER_SM
#define ABIP_MIN_BLOCKS_PER_SM 2
#endif
constexpr int kBlock = ABIP_BLOCK;
constexpr int kWarps = kBlock / 32;
constexpr int kMaxRed = 24;  // max scalars reduced between two grid barriers

// ---------------------------------------------------------------------------------------------------------
// Streaming loads for the matrix arrays: read-only path, do not allocate in L1 (keeps L1 for the gathered
// vector), default L2 policy (A and A' together are ~L2-sized at cfg2, so we want them retained in L2).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream4(const int* p) {  // 16-byte aligned
    int4 v;
    asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {  // 16-byte aligned
    double2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// CSR matrix + the SpMV variant chosen from its row-length statistics (host: choose_spmv_plan()).
struct Csr {
    const int* ptr;        // [nrows+1]
    const int* idx;        // [nnz] column indices
    const double* val;     // [nnz]
    int nrows;
    int lanes_log2;        // vector-per-row width L = 1<<lanes_log2 lanes (1..32) for rows <= long_thresh
    int long_thresh;       // rows longer than this go to the warp-per-row 128-bit path
    const int* long_rows;  // [n_long] their indices
    int n_long;
};

// Deterministic grid-wide reduction helper (see file header).  partials is double-buffered so that a fast
// block starting reduction t+1 can never overwrite values a slow block is still reading for reduction t.
struct Reducer {
    double* partials;  // [2][kMaxRed][G]
    double* sm;        // shared scratch [kMaxRed * kWarps]
    int G;
    int parity;

    // adds this block's contribution for slots [slot0, slot0+K); several calls (distinct slots) may precede one
    // grid barrier + finish()
    template <int K>
    __device__ __forceinline__ void block_store(double (&v)[K], int slot0 = 0) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
            if (lane == 0) sm[k * kWarps + w] = x;
        }
        __syncthreads();
        if (threadIdx.x < K) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) s += sm[threadIdx.x * kWarps + i];
            partials[(parity * kMaxRed + slot0 + threadIdx.x) * G + blockIdx.x] = s;
        }
        __syncthreads();
        // the grid barrier that follows orders these writes before finish()
    }

    // call after the grid barrier; every block computes the same totals in the same order
    template <int K>
    __device__ __forceinline__ void finish(double (&out)[K]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int k = w; k < K; k += kWarps) {
            const double* src = partials + (parity * kMaxRed + k) * G;
            double s = 0.0;
            for (int i = lane; i < G; i += 32) s += __ldcg(src + i);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
            if (lane == 0) sm[k] = s;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = sm[k];
        __syncthreads();
        parity ^= 1;
    }
};

// ---------------------------------------------------------------------------------------------------------
// K1: CSR SpMV phase.  fn(row, dot) is called once per row by one lane with dot = A[row,:] * x.
//   * rows of length <= long_thresh: vector-per-row, L = 2^lanes_log2 lanes per row, warp-coalesced scalar
//     loads (adjacent sub-warps read adjacent rows, so a warp streams one contiguous span of val/idx);
//   * longer rows: warp-per-row with 128-bit loads (int4 indices, 2 x double2 values) after an alignment peel.
// x may have been written earlier in the same kernel (ordinary coherent loads; L1 is invalidated by the grid
// barrier's fence).  Replaces _accum_by_Atrans (reference linsys/common.c:598-639) for both A and A'.
// ---------------------------------------------------------------------------------------------------------
template <class RowFn>
__device__ __forceinline__ void spmv_rows(const Csr& A, const double* x, RowFn fn) {
    const int lg = A.lanes_log2;
    const int L = 1 << lg;
    const int lane = threadIdx.x & 31;
    const int sl = lane & (L - 1);
    const int sub = lane >> lg;
    const int rpw = 32 >> lg;
    const int gwarp = blockIdx.x * kWarps + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * kWarps;
    const int nrows = A.nrows;
    for (int base = gwarp * rpw; base < nrows; base += nwarps * rpw) {
        const int row = base + sub;
        int s = 0, e = 0;
        bool ok = row < nrows;
        if (ok) {
            s = __ldg(A.ptr + row);
            e = __ldg(A.ptr + row + 1);
            if (e - s > A.long_thresh) { ok = false; e = s; }
        }
        double acc = 0.0;
        for (int k = s + sl; k < e; k += L) acc = fma(ld_stream(A.val + k), x[ld_stream(A.idx + k)], acc);
        for (int off = L >> 1; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off, L);
        if (ok && sl == 0) fn(row, acc);
    }
    for (int li = gwarp; li < A.n_long; li += nwarps) {
        const int row = __ldg(A.long_rows + li);
        const int s = __ldg(A.ptr + row), e = __ldg(A.ptr + row + 1);
        double acc = 0.0;
        int s4 = (s + 3) & ~3;
        if (s4 > e) s4 = e;
        if (s + lane < s4) acc = ld_stream(A.val + s + lane) * x[ld_stream(A.idx + s + lane)];
        const int e4 = s4 + ((e - s4) & ~3);
        for (int k = s4 + lane * 4; k < e4; k += 128) {
            const int4 c = ld_stream4(A.idx + k);
            const double2 v0 = ld_stream2(A.val + k), v1 = ld_stream2(A.val + k + 2);
            const double x0 = x[c.x], x1 = x[c.y], x2 = x[c.z], x3 = x[c.w];
            acc = fma(v0.x, x0, acc);
            acc = fma(v0.y, x1, acc);
            acc = fma(v1.x, x2, acc);
            acc = fma(v1.y, x3, acc);
        }
        if (e4 + lane < e) acc = fma(ld_stream(A.val + e4 + lane), x[ld_stream(A.idx + e4 + lane)], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) fn(row, acc);
    }
}

// Constant problem data + PCG workspace (kernel parameter, passed by value).
struct LpCtx {
    int m, n;
    Csr A;   // CSR(A): m rows.   Reference keeps it as "At" in CSC (linsys/indirect.c:81-139)
    Csr AT;  // CSR(A'): n rows == the caller's CSC arrays of A
    const double* M;  // 1/diag(AA')   (indirect.c:36-79; no rho_y term -- parity trap 1)
    const double* D;  // row scaling or nullptr
    const double* E;  // col scaling or nullptr
    const double* b;  // scaled b [m]
    const double* c;  // scaled c [n]
    const double* h;  // [-b; c] [m+n]
    const double* g;  // K^-1 h with g_x flipped [m+n]
    double rho_y, alpha, cg_rate, g_th;
    double *p, *r, *Gp, *tmp;  // PCG: direction, residual, G p [m]; A'p scratch [n]
    double* partials;
    double* sc;  // scalar block [ABIPGPU_SC_COUNT]
};

struct SolveOut {
    int its;
    double tol, res;
};

#define GRID_STRIDE(i, N) \
    for (int i = blockIdx.x * kBlock + threadIdx.x, _gs = gridDim.x * kBlock; i < (N); i += _gs)

// ---------------------------------------------------------------------------------------------------------
// solve_lin_sys on device (reference linsys/indirect.c:393-434 incl. pcg :321-391 and mat_vec :205-220).
//   b: [m+n] right-hand side, overwritten by the solution;  s: warm start [>= m] or nullptr.
//   EPI: also reduce hdot = sol[0:m+n] . h (epilogue of project_lin_sys, src/abip.c:560); the caller must
//   grid.sync() and R.finish<1>() to obtain it.
// Barriers per solve: 2 + 4 per CG iteration (+1 by the caller).
// ---------------------------------------------------------------------------------------------------------
template <bool EPI>
__device__ __forceinline__ void dev_solve_lin_sys(const LpCtx& c, Reducer& R, cg::grid_group& grid, double* b,
                                                  const double* s, long iter, SolveOut& out) {
    const int m = c.m;
    double* by = b;
    double* bx = b + m;
    // S1: by += A bx, accumulating |by|^2 of the *incoming* by for the tolerance (indirect.c:406-409, trap 2);
    //     independent of that, tmp = A' s for the warm-start residual.
    double a1[1] = {0.0};
    spmv_rows(c.A, bx, [&](int row, double a) {
        const double o = by[row];
        a1[0] = fma(o, o, a1[0]);
        by[row] = o + a;
    });
    if (s) spmv_rows(c.AT, s, [&](int row, double a) { c.tmp[row] = a; });
    R.block_store<1>(a1);
    grid.sync();
    R.finish<1>(a1);
    double tol = sqrt(a1[0]) * (iter < 0 ? 1e-9 : 0.1 / pow((double)iter + 1.0, c.cg_rate));
    tol = fmax(fmax(tol, 1e-7), 1e-9);
    // S2: r = by - (rho s + A tmp), x = s (stored in by), z = M r, p = z   (indirect.c:343-365)
    double a2[2] = {0.0, 0.0};
    if (s) {
        spmv_rows(c.A, c.tmp, [&](int row, double a) {
            const double si = s[row];
            const double ri = by[row] - fma(c.rho_y, si, a);
            const double zi = __ldg(c.M + row) * ri;
            c.r[row] = ri;
            by[row] = si;
            c.p[row] = zi;
            a2[0] = fma(ri, ri, a2[0]);
            a2[1] = fma(zi, ri, a2[1]);
        });
    } else {
        GRID_STRIDE(i, m) {
            const double ri = by[i];
            const double zi = __ldg(c.M + i) * ri;
            c.r[i] = ri;
            by[i] = 0.0;
            c.p[i] = zi;
            a2[0] = fma(ri, ri, a2[0]);
            a2[1] = fma(zi, ri, a2[1]);
        }
    }
    R.block_store<2>(a2);
    grid.sync();
    R.finish<2>(a2);
    double rn = sqrt(a2[0]);
    double ipzr = a2[1];
    int its = 0;
    if (!(rn < fmin(tol, 1e-18))) {
        for (int it = 0; it < m; ++it) {
            // L1: tmp = A' p
            spmv_rows(c.AT, c.p, [&](int row, double a) { c.tmp[row] = a; });
            grid.sync();
            // L2: Gp = A tmp + rho p ; p.Gp
            double d1[1] = {0.0};
            spmv_rows(c.A, c.tmp, [&](int row, double a) {
                const double pi = c.p[row];
                const double gp = fma(c.rho_y, pi, a);
                c.Gp[row] = gp;
                d1[0] = fma(pi, gp, d1[0]);
            });
            R.block_store<1>(d1);
            grid.sync();
            R.finish<1>(d1);
            const double alpha = ipzr / d1[0];
            // L3: x += alpha p ; r -= alpha Gp ; |r|^2 ; (M r).r
            double d2[2] = {0.0, 0.0};
            GRID_STRIDE(i, m) {
                by[i] = fma(alpha, c.p[i], by[i]);
                const double ri = fma(-alpha, c.Gp[i], c.r[i]);
                c.r[i] = ri;
                const double zi = __ldg(c.M + i) * ri;
                d2[0] = fma(ri, ri, d2[0]);
                d2[1] = fma(zi, ri, d2[1]);
            }
            R.block_store<2>(d2);
            grid.sync();
            R.finish<2>(d2);
            its = it + 1;
            rn = sqrt(d2[0]);
            if (rn < tol) break;
            const double beta = d2[1] / ipzr;
            ipzr = d2[1];
            // L4: p = beta p + M r
            GRID_STRIDE(i, m) c.p[i] = fma(beta, c.p[i], __ldg(c.M + i) * c.r[i]);
            grid.sync();
        }
    }
    // S4: bx = -bx + A' by   (indirect.c:419-420)
    double a3[1] = {0.0};
    spmv_rows(c.AT, by, [&](int row, double a) {
        const double nv = a - bx[row];
        bx[row] = nv;
        if (EPI) a3[0] = fma(nv, __ldg(c.h + m + row), a3[0]);
    });
    if (EPI) {
        GRID_STRIDE(i, m) a3[0] = fma(by[i], __ldg(c.h + i), a3[0]);
        R.block_store<1>(a3);
    }
    out.its = its;
    out.tol = tol;
    out.res = rn;
}

// ---------------------------------------------------------------------------------------------------------
// Right-hand side of project_lin_sys (reference src/abip.c:551-558, src/adaptive.c:91-96):
//   ut = u + v; ut[0:m] *= rho_y; ut[0:l-1] -= ut[l-1] h; ut[0:l-1] -= h (ut.g)/(g_th+1); ut[m:l-1] *= -1
// Optionally records u_prev = u on the (x,tau) tail (abip.c:2133; only the tail is ever read back).
// Ends with a grid barrier.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dev_build_rhs(const LpCtx& c, Reducer& R, cg::grid_group& grid, const double* u,
                                              const double* v, double* ut, double* u_prev_out) {
    const int m = c.m, lm1 = c.m + c.n;
    const double tt = u[lm1] + v[lm1];
    double a[1] = {0.0};
    GRID_STRIDE(i, lm1) {
        const double ui = u[i];
        double w = ui + v[i];
        if (i < m) w *= c.rho_y;
        else if (u_prev_out) u_prev_out[i] = ui;
        w = fma(-tt, __ldg(c.h + i), w);
        ut[i] = w;
        a[0] = fma(w, __ldg(c.g + i), a[0]);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ut[lm1] = tt;
        if (u_prev_out) u_prev_out[lm1] = u[lm1];
    }
    R.block_store<1>(a);
    grid.sync();
    R.finish<1>(a);
    const double coef = -a[0] / (c.g_th + 1.0);
    GRID_STRIDE(i, lm1) {
        double w = fma(coef, __ldg(c.h + i), ut[i]);
        if (i >= m) w = -w;
        ut[i] = w;
    }
    grid.sync();
}

// barrier proximal step on one coordinate (src/abip.c:742-746)
__device__ __forceinline__ double barrier_prox(double t, double lam) {
    const double hlf = t / 2;
    return hlf + sqrt(fma(hlf, hlf, lam));
}
