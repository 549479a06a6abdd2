// Round-2 micro-benchmark: one PCG iteration of ABIP-LP (A'p pass, A pass + dots, vector update) on a cfg2-shaped
// multicommodity-flow matrix, with the SJDS layout + locality ordering of abip_b200/csrc/sjds_host.h, in a persistent
// cooperative grid.  Variants: natural / reordered matrix, cg::grid.sync / counter barrier, 4-barrier classic CG phase
// structure / 3-barrier fused structure, unroll depth.  `--cpu` only prints the gather model (no GPU needed).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -I abip_b200/csrc -I tools/ubench tools/ubench/sjds_bench.cu -o build/sjds_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <random>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include "sjds_host.h"
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

#ifndef BLOCK
#define BLOCK 512
#endif
#ifndef BPSM
#define BPSM 2
#endif
constexpr int kBlock = BLOCK, kWarps = BLOCK / 32;

#include "sjds_device.cuh"
typedef Sjds Mat;
constexpr size_t kRingBytes = RingLayout<kWarps>::bytes;

// ---- grid barrier: arrive counter + generation flag (release / acquire at gpu scope) -----------------------------
struct GBar { unsigned* count; unsigned* gen; };
__device__ __forceinline__ void gbar_sync(const GBar& b, unsigned& my_gen) {
    __syncthreads();
    if (threadIdx.x == 0) {
        my_gen += 1;
        unsigned prev;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(prev) : "l"(b.count) : "memory");
        if (prev == gridDim.x - 1) {
            *(volatile unsigned*)b.count = 0;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(b.gen), "r"(my_gen) : "memory");
        } else {
            unsigned g;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(g) : "l"(b.gen) : "memory"); } while (g < my_gen);
        }
        __threadfence();  // L1 of this SM must not serve stale lines of vectors written by other SMs
    }
    __syncthreads();
}

constexpr int kMaxRed = 8;
struct Red {
    double* partials; double* sm; int parity;
    template <int K> __device__ __forceinline__ void store(double (&v)[K]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = v[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
            if (lane == 0) sm[k * kWarps + w] = x;
        }
        __syncthreads();
        if (threadIdx.x < K) {
            double s = 0.0;
            for (int i = 0; i < kWarps; ++i) s += sm[threadIdx.x * kWarps + i];
            partials[(parity * kMaxRed + threadIdx.x) * gridDim.x + blockIdx.x] = s;
        }
    }
    template <int K> __device__ __forceinline__ void finish(double (&out)[K]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        const int G = gridDim.x;
        for (int k = w; k < K; k += kWarps) {
            const double* src = partials + (parity * kMaxRed + k) * G;
            double s = 0.0;
            for (int i = lane; i < G; i += 32) s += __ldcg(src + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) sm[k] = s;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) out[k] = sm[k];
        __syncthreads();
        parity ^= 1;
    }
};

struct Ctx {
    Mat A, AT; int m, n;
    double *p, *r, *Gp, *x, *tmp; const double* M;
    double* partials; GBar bar; double rho; double* out;
};

#define GRID_STRIDE(i, N) for (int i = blockIdx.x * kBlock + threadIdx.x, _gs = gridDim.x * kBlock; i < (N); i += _gs)

// mode bit 0: counter barrier (else cg grid.sync); bit 1: fused 3-barrier CG iteration (else classic 4 barriers)
// bit 2: passes only (no vector phases), bit 3: A' pass only, bit 4: A pass only, bit 5: barriers only
template <int U>
__global__ void __launch_bounds__(kBlock, BPSM) k_cg(Ctx c, int iters, int mode) {
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpRing rg = make_ring<kWarps>(smem_raw);
    __shared__ double red_sm[kMaxRed * kWarps];
    Red R{c.partials, red_sm, 0};
    unsigned gen = *(volatile unsigned*)c.bar.gen;
    auto bar = [&]() { if (mode & 1) gbar_sync(c.bar, gen); else grid.sync(); };
    const int m = c.m;
    double ipzr = 1.0;
    for (int it = 0; it < iters; ++it) {
        if (mode & 32) { bar(); bar(); bar(); continue; }
        if (!(mode & 16)) sjds_rows<kBlock>(c.AT, c.p, rg, (mode & 8) ? &c.AT : &c.A, blockIdx.x, [&](int row, double a) { c.tmp[row] = a; });
        if (mode & 8) { bar(); continue; }
        bar();
        if (mode & 2) {
            double d[7] = {0, 0, 0, 0, 0, 0, 0};
            sjds_rows<kBlock>(c.A, c.tmp, rg, (mode & 16) ? &c.A : &c.AT, blockIdx.x, [&](int row, double a) {
                const double pi = c.p[row], ri = c.r[row], Mi = __ldg(c.M + row);
                const double gp = fma(c.rho, pi, a);
                c.Gp[row] = gp;
                const double zi = Mi * ri, mg = Mi * gp;
                d[0] = fma(pi, gp, d[0]); d[1] = fma(zi, gp, d[1]); d[2] = fma(mg, gp, d[2]);
                d[3] = fma(ri, gp, d[3]); d[4] = fma(gp, gp, d[4]); d[5] = fma(zi, ri, d[5]); d[6] = fma(ri, ri, d[6]);
            });
            if (mode & 4) { bar(); continue; }
            R.store<7>(d);
            bar();
            R.finish<7>(d);
            const double alpha = 1e-300 * (d[5] / d[0]);
            const double a2 = d[5] - 2 * alpha * d[1] + alpha * alpha * d[2];
            const double beta = 1e-300 * (a2 / d[5]) + 0.5;
            GRID_STRIDE(i, m) {
                const double pi = c.p[i];
                c.x[i] = fma(alpha, pi, c.x[i]);
                const double ri = fma(-alpha, c.Gp[i], c.r[i]);
                c.r[i] = ri;
                c.p[i] = fma(beta, pi, __ldg(c.M + i) * ri);
            }
            bar();
        } else {
            double d1[1] = {0};
            sjds_rows<kBlock>(c.A, c.tmp, rg, (mode & 16) ? &c.A : &c.AT, blockIdx.x, [&](int row, double a) {
                const double pi = c.p[row];
                const double gp = fma(c.rho, pi, a);
                c.Gp[row] = gp;
                d1[0] = fma(pi, gp, d1[0]);
            });
            if (mode & 4) { bar(); continue; }
            R.store<1>(d1);
            bar();
            R.finish<1>(d1);
            const double alpha = 1e-300 * (ipzr / d1[0]);
            double d2[2] = {0, 0};
            GRID_STRIDE(i, m) {
                c.x[i] = fma(alpha, c.p[i], c.x[i]);
                const double ri = fma(-alpha, c.Gp[i], c.r[i]);
                c.r[i] = ri;
                const double zi = __ldg(c.M + i) * ri;
                d2[0] = fma(ri, ri, d2[0]); d2[1] = fma(zi, ri, d2[1]);
            }
            R.store<2>(d2);
            bar();
            R.finish<2>(d2);
            const double beta = 1e-300 * (d2[1] / ipzr) + 0.5;
            ipzr = d2[1];
            GRID_STRIDE(i, m) c.p[i] = fma(beta, c.p[i], __ldg(c.M + i) * c.r[i]);
            bar();
        }
    }
    rg.drain();
    if (blockIdx.x == 0 && threadIdx.x == 0) c.out[0] = ipzr;
}

__global__ void __launch_bounds__(kBlock, BPSM) k_spmv(Mat A, const double* x, double* y) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WarpRing rg = make_ring<kWarps>(smem_raw);
    sjds_rows<kBlock>(A, x, rg, nullptr, blockIdx.x, [&](int row, double a) { y[row] = a; });
    rg.drain();
}

// ---- host ------------------------------------------------------------------------------------------------------
struct Coo { std::vector<int> r, c; std::vector<double> v; };
static void to_csr(int nrows, const Coo& C, bool by_col, std::vector<int>* ptr, std::vector<int>* idx, std::vector<double>* val) {
    const std::vector<int>& R = by_col ? C.c : C.r;
    const std::vector<int>& Cc = by_col ? C.r : C.c;
    const size_t nnz = R.size();
    ptr->assign(nrows + 1, 0);
    for (size_t k = 0; k < nnz; ++k) (*ptr)[R[k] + 1]++;
    for (int i = 0; i < nrows; ++i) (*ptr)[i + 1] += (*ptr)[i];
    idx->resize(nnz); val->resize(nnz);
    std::vector<int> fill(ptr->begin(), ptr->end() - 1);
    for (size_t k = 0; k < nnz; ++k) { const int q = fill[R[k]]++; (*idx)[q] = Cc[k]; (*val)[q] = C.v[k]; }
    // sort each row by column
    std::vector<std::pair<int, double>> tmp;
    for (int i = 0; i < nrows; ++i) {
        const int a = (*ptr)[i], b = (*ptr)[i + 1];
        tmp.resize(b - a);
        for (int k = a; k < b; ++k) tmp[k - a] = {(*idx)[k], (*val)[k]};
        std::sort(tmp.begin(), tmp.end());
        for (int k = a; k < b; ++k) { (*idx)[k] = tmp[k - a].first; (*val)[k] = tmp[k - a].second; }
    }
}

static Coo gen_mcf(int K, int V, int E, int R, int w, int* m, int* n, double scale_side) {
    std::mt19937_64 rng(12345);
    Coo C;
    std::vector<int> tail(E), head(E);
    for (int e = 0; e < E; ++e) { tail[e] = rng() % V; head[e] = (tail[e] + 1 + rng() % (V - 1)) % V; }
    std::normal_distribution<double> nd(0, 1);
    const int nflow = K * E;
    *n = nflow + E;
    *m = K * V + E + R;
    for (int k = 0; k < K; ++k)
        for (int e = 0; e < E; ++e) {
            const int col = k * E + e;
            C.r.push_back(k * V + tail[e]); C.c.push_back(col); C.v.push_back(1.0 + 0.01 * nd(rng));
            C.r.push_back(k * V + head[e]); C.c.push_back(col); C.v.push_back(-1.0 + 0.01 * nd(rng));
            C.r.push_back(K * V + e); C.c.push_back(col); C.v.push_back(1.0);
        }
    for (int e = 0; e < E; ++e) { C.r.push_back(K * V + e); C.c.push_back(nflow + e); C.v.push_back(1.0); }
    for (int r = 0; r < R; ++r) {
        std::vector<int> cs(w);
        for (int q = 0; q < w; ++q) cs[q] = rng() % nflow;
        std::sort(cs.begin(), cs.end());
        cs.erase(std::unique(cs.begin(), cs.end()), cs.end());
        for (int cc : cs) { C.r.push_back(K * V + E + r); C.c.push_back(cc); C.v.push_back(scale_side * nd(rng)); }
    }
    return C;
}

struct DevMat { Mat view; std::vector<void*> allocs; };
template <class T> static T* up(const std::vector<T>& v, std::vector<void*>& allocs, size_t pad = 64) {
    T* d; CK(cudaMalloc(&d, (v.size() + pad) * sizeof(T))); CK(cudaMemset(d, 0, (v.size() + pad) * sizeof(T)));
    if (!v.empty()) CK(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    allocs.push_back(d); return d;
}
static DevMat upload(const sjds::Host& H, int slot) {
    DevMat D;
    std::vector<int4> ch(H.chunk.size()), lr(H.long_rows.size());
    for (size_t i = 0; i < ch.size(); ++i) ch[i] = make_int4(H.chunk[i].start, H.chunk[i].count, H.chunk[i].row0, H.chunk[i].info);
    for (size_t i = 0; i < lr.size(); ++i) lr[i] = make_int4(H.long_rows[i].row, H.long_rows[i].slot0, H.long_rows[i].npieces, 0);
    std::vector<double> lp(H.n_pieces + 8, 0.0);
    D.view = Mat{up(H.val, D.allocs, kCH + 8), up(H.idx, D.allocs, kCH + 8), up(H.meta, D.allocs), up(ch, D.allocs), up(H.warp_chunk, D.allocs),
                 H.n_long ? up(lr, D.allocs) : nullptr, up(H.cta_long, D.allocs), up(lp, D.allocs), H.nrows, slot};
    return D;
}

static void cpu_spmv(int nrows, const std::vector<int>& ptr, const std::vector<int>& idx, const std::vector<double>& val,
                     const std::vector<double>& x, std::vector<double>* y) {
    y->assign(nrows, 0.0);
    for (int i = 0; i < nrows; ++i) { double a = 0; for (int k = ptr[i]; k < ptr[i + 1]; ++k) a += val[k] * x[idx[k]]; (*y)[i] = a; }
}

int main(int argc, char** argv) {
    bool cpu_only = false; double scale = 1.0;
    for (int i = 1; i < argc; ++i) { if (!strcmp(argv[i], "--cpu")) cpu_only = true; else scale = atof(argv[i]); }
    const int K = 20, V = (int)(7500 * scale), E = (int)(47600 * scale), R = (int)(2400 * scale), w = (int)(875 * std::min(1.0, scale * 4));
    int m, n;
    Coo C = gen_mcf(K, V, E, R, w, &m, &n, 1.0);
    std::vector<int> a_ptr, a_idx, at_ptr, at_idx; std::vector<double> a_val, at_val;
    to_csr(m, C, false, &a_ptr, &a_idx, &a_val);
    to_csr(n, C, true, &at_ptr, &at_idx, &at_val);
    printf("MCF m=%d n=%d nnz=%d\n", m, n, a_ptr[m]);
    int G = 148 * BPSM;
    if (!cpu_only) { CK(cudaFuncSetAttribute((const void*)k_cg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingBytes)); CK(cudaFuncSetAttribute((const void*)k_spmv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingBytes)); cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); G = pr.multiProcessorCount * BPSM; printf("%s, %d SMs, grid %d x %d, ring %zu B/CTA (CH=%d NS=%d GU=%d)\n", pr.name, pr.multiProcessorCount, G, kBlock, kRingBytes, kCH, kNS, ABIP_GU); }

    for (int reorder = 0; reorder < 2; ++reorder) {
        std::vector<int> rn2o(m), cn2o(n);
        for (int i = 0; i < m; ++i) rn2o[i] = i;
        for (int j = 0; j < n; ++j) cn2o[j] = j;
        if (reorder) sjds::locality_order(m, n, a_ptr, a_idx, at_ptr, at_idx, sjds::kLongRow, &rn2o, &cn2o, [](long nn, auto fn) { fn(0L, nn, 0); });
        std::vector<int> ro2n(m), co2n(n);
        for (int i = 0; i < m; ++i) ro2n[rn2o[i]] = i;
        for (int j = 0; j < n; ++j) co2n[cn2o[j]] = j;
        std::vector<int> pa_ptr, pa_idx, pat_ptr, pat_idx; std::vector<double> pa_val, pat_val;
        sjds::permute_csr(m, a_ptr, a_idx, a_val, rn2o, co2n, &pa_ptr, &pa_idx, &pa_val);
        sjds::permute_csr(n, at_ptr, at_idx, at_val, cn2o, ro2n, &pat_ptr, &pat_idx, &pat_val);
        sjds::Host HA, HAT;
        sjds::build(m, n, pa_ptr, pa_idx, pa_val, G, kWarps, &HA);
        sjds::build(n, m, pat_ptr, pat_idx, pat_val, G, kWarps, &HAT);
        const sjds::GatherStats ga = sjds::gather_stats(HA), gat = sjds::gather_stats(HAT);
        printf("[%s] A : %d slices, %d long rows, %zu chunks, stored %zu (nnz %ld); gather: %.0f req, %.2f lanes, %.2f lines, %.2f sectors per request\n",
               reorder ? "reordered" : "natural", HA.nslices, HA.n_long, HA.chunk.size(), HA.val.size(), HA.nnz, ga.requests, ga.lanes / ga.requests, ga.lines / ga.requests, ga.sectors / ga.requests);
        printf("[%s] A': %d slices, %d long rows, %zu chunks, stored %zu; gather: %.0f req, %.2f lanes, %.2f lines, %.2f sectors per request\n",
               reorder ? "reordered" : "natural", HAT.nslices, HAT.n_long, HAT.chunk.size(), HAT.val.size(), gat.requests, gat.lanes / gat.requests, gat.lines / gat.requests, gat.sectors / gat.requests);
        if (cpu_only) continue;

        DevMat DA = upload(HA, 1), DAT = upload(HAT, 2);
        std::vector<void*> al;
        std::vector<double> hp(m), hM(m, 0.5), hx(n);
        std::mt19937_64 rng(7);
        std::normal_distribution<double> nd(0, 1);
        for (auto& v : hp) v = nd(rng);
        for (auto& v : hx) v = nd(rng);
        Ctx c;
        c.A = DA.view; c.AT = DAT.view; c.m = m; c.n = n;
        c.p = up(hp, al); c.r = up(hp, al); c.Gp = up(hp, al); c.x = up(hp, al); c.M = up(hM, al);
        c.tmp = up(hx, al);
        std::vector<double> z(2 * kMaxRed * G + 64, 0.0);
        c.partials = up(z, al);
        std::vector<unsigned> bz(64, 0);
        unsigned* dbar = up(bz, al);
        c.bar = GBar{dbar, dbar + 32};
        c.rho = 1e-3;
        c.out = up(z, al);
        // correctness of both passes against the CPU
        {
            std::vector<double> yref, y(std::max(m, n));
            double* dy; CK(cudaMalloc(&dy, sizeof(double) * std::max(m, n)));
            cpu_spmv(m, pa_ptr, pa_idx, pa_val, hx, &yref);
            k_spmv<<<G, kBlock, kRingBytes>>>(c.A, c.tmp, dy); CK(cudaGetLastError());
            CK(cudaMemcpy(y.data(), dy, sizeof(double) * m, cudaMemcpyDeviceToHost));
            double err = 0, nrm = 0;
            for (int i = 0; i < m; ++i) { err = std::max(err, fabs(y[i] - yref[i])); nrm = std::max(nrm, fabs(yref[i])); }
            printf("  A  pass max err %.3e (max |y| %.3e)\n", err, nrm);
            cpu_spmv(n, pat_ptr, pat_idx, pat_val, hp, &yref);
            k_spmv<<<G, kBlock, kRingBytes>>>(c.AT, c.p, dy); CK(cudaGetLastError());
            CK(cudaMemcpy(y.data(), dy, sizeof(double) * n, cudaMemcpyDeviceToHost));
            err = 0; nrm = 0;
            for (int i = 0; i < n; ++i) { err = std::max(err, fabs(y[i] - yref[i])); nrm = std::max(nrm, fabs(yref[i])); }
            printf("  A' pass max err %.3e (max |y| %.3e)\n", err, nrm);
            CK(cudaFree(dy));
        }
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        auto run = [&](const char* name, const void* kern, int mode, int iters) {
            void* args[] = {(void*)&c, (void*)&iters, (void*)&mode};
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                CK(cudaLaunchCooperativeKernel(kern, dim3(G), dim3(kBlock), args, kRingBytes, 0));
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (rep) best = std::min(best, ms);
            }
            printf("  %-58s %8.2f us / iteration\n", name, 1e3 * best / iters);
        };
        {
            double* dy; CK(cudaMalloc(&dy, sizeof(double) * std::max(m, n)));
            float ta = 0, tat = 0;
            const int reps = 30;
            for (int rep = -3; rep < reps; ++rep) {
                float ms;
                CK(cudaEventRecord(e0)); k_spmv<<<G, kBlock, kRingBytes>>>(c.AT, c.p, dy); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep >= 0) tat += ms;
                CK(cudaEventRecord(e0)); k_spmv<<<G, kBlock, kRingBytes>>>(c.A, c.tmp, dy); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep >= 0) ta += ms;
            }
            printf("  alternating stand-alone launches   A' %7.2f us   A %7.2f us\n", 1e3 * tat / reps, 1e3 * ta / reps);
            CK(cudaFree(dy));
        }
        const int IT = 200;
        run("barriers only x3, cg::grid.sync", (const void*)k_cg<4>, 32, IT);
        run("A' pass + barrier", (const void*)k_cg<4>, 8, IT);
        run("A  pass (1 dot) + 2 barriers", (const void*)k_cg<4>, 16 | 4, IT);
        run("A  pass (7 dots) + 2 barriers", (const void*)k_cg<4>, 16 | 4 | 2, IT);
        run("both passes + 2 barriers", (const void*)k_cg<4>, 4, IT);
        run("CG iteration classic 4 barriers, cg sync", (const void*)k_cg<4>, 0, IT);
        run("CG iteration classic 4 barriers, counter barrier", (const void*)k_cg<4>, 1, IT);
        run("CG iteration fused 3 barriers, cg sync", (const void*)k_cg<4>, 2, IT);
        run("CG iteration fused 3 barriers, counter barrier", (const void*)k_cg<4>, 3, IT);
        for (void* q : DA.allocs) cudaFree(q);
        for (void* q : DAT.allocs) cudaFree(q);
        for (void* q : al) cudaFree(q);
    }
    return 0;
}
