// Micro-benchmark: how fast can a persistent grid stream CSR chunks (2 KB of values + 1 KB of indices per chunk)
// from HBM into shared memory with the mechanisms available on sm_100a?  Guides the SpMV staging design.
//   mode 0: per-warp cp.async (16 B per lane), S stages per warp
//   mode 1: per-warp TMA bulk copies (cp.async.bulk + mbarrier), S stages per warp
//   mode 2: LDG.128 into registers (no shared memory)
// Each "chunk" is consumed by a token reduction (one LDS.128 per lane) so that nothing is optimised away.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int VALB = 2048, IDXB = 1024, STAGE = VALB + IDXB;
__device__ __forceinline__ unsigned s32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int S>
__global__ void __launch_bounds__(1024, 1) k_cpasync(const double* val, const int* idx, long nchunks, int warps, double* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= warps) return;
    const long W = (long)gridDim.x * warps, gw = (long)blockIdx.x * warps + w;
    const long c0 = nchunks * gw / W, c1 = nchunks * (gw + 1) / W;
    unsigned char* base = sm + (size_t)w * S * STAGE;
    double acc = 0;
    auto issue = [&](long c, int st) {
        const unsigned d = s32(base + st * STAGE) + 16 * lane;
        const char* v = (const char*)(val + c * 256) + 16 * lane;
        const char* i = (const char*)(idx + c * 256) + 16 * lane;
#pragma unroll
        for (int j = 0; j < 4; ++j) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 512 * j), "l"(v + 512 * j) : "memory");
#pragma unroll
        for (int j = 0; j < 2; ++j) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + VALB + 512 * j), "l"(i + 512 * j) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    long ci = c0;
    for (int s = 0; s < S - 1 && ci < c1; ++s, ++ci) issue(ci, s);
    int st = 0;
    for (long c = c0; c < c1; ++c) {
        if (ci < c1) { issue(ci, (st + S - 1) % S); ++ci; } else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");
        __syncwarp();
        const double2 t = reinterpret_cast<const double2*>(base + st * STAGE)[lane];
        acc += t.x + t.y;
        __syncwarp();
        st = (st + 1) % S;
    }
    if (acc == 1.2345e-300) out[0] = acc;
}

template <int S>
__global__ void __launch_bounds__(1024, 1) k_tma(const double* val, const int* idx, long nchunks, int warps, double* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bars[32 * S];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= warps) return;
    const long W = (long)gridDim.x * warps, gw = (long)blockIdx.x * warps + w;
    const long c0 = nchunks * gw / W, c1 = nchunks * (gw + 1) / W;
    unsigned char* base = sm + (size_t)w * S * STAGE;
    if (lane == 0)
        for (int s = 0; s < S; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[w * S + s])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    double acc = 0;
    auto issue = [&](long c, int st) {
        if (lane == 0) {
            const unsigned b = s32(&bars[w * S + st]), d = s32(base + st * STAGE);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(STAGE) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d), "l"(val + c * 256), "r"(VALB), "r"(b) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(d + VALB), "l"(idx + c * 256), "r"(IDXB), "r"(b) : "memory");
        }
    };
    long ci = c0;
    for (int s = 0; s < S - 1 && ci < c1; ++s, ++ci) issue(ci, s);
    int st = 0; unsigned ph = 0;
    for (long c = c0; c < c1; ++c) {
        if (ci < c1) { issue(ci, (st + S - 1) % S); ++ci; }
        const unsigned b = s32(&bars[w * S + st]);
        unsigned ok = 0;
        while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(b), "r"((ph >> st) & 1u) : "memory");
        const double2 t = reinterpret_cast<const double2*>(base + st * STAGE)[lane];
        acc += t.x + t.y;
        __syncwarp();
        ph ^= 1u << st;
        st = (st + 1) % S;
    }
    if (acc == 1.2345e-300) out[0] = acc;
}

__global__ void __launch_bounds__(1024, 1) k_ldg(const double* val, const int* idx, long nchunks, int warps, double* out) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (w >= warps) return;
    const long W = (long)gridDim.x * warps, gw = (long)blockIdx.x * warps + w;
    const long c0 = nchunks * gw / W, c1 = nchunks * (gw + 1) / W;
    double acc = 0;
    for (long c = c0; c < c1; ++c) {
        const double2* v = reinterpret_cast<const double2*>(val + c * 256) + lane;
        const int4* i = reinterpret_cast<const int4*>(idx + c * 256) + lane;
        double2 a0 = __ldcs(v), a1 = __ldcs(v + 32), a2 = __ldcs(v + 64), a3 = __ldcs(v + 96);
        int4 i0 = __ldcs(i), i1 = __ldcs(i + 32);
        acc += a0.x + a1.y + a2.x + a3.y + (double)(i0.x + i1.w);
    }
    if (acc == 1.2345e-300) out[0] = acc;
}

int main() {
    const long nchunks = 40000;  // 2 matrices x 20000 chunks x 3 KB = 120 MB
    double* val; int* idx; double* out;
    CK(cudaMalloc(&val, nchunks * 256 * 8)); CK(cudaMalloc(&idx, nchunks * 256 * 4)); CK(cudaMalloc(&out, 8));
    CK(cudaMemset(val, 0, nchunks * 256 * 8)); CK(cudaMemset(idx, 0, nchunks * 256 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch) {
        for (int i = 0; i < 3; ++i) launch();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int R = 20;
        for (int i = 0; i < R; ++i) launch();
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-40s %8.1f us/pass  %7.1f GB/s\n", name, ms / R * 1e3, nchunks * 3072.0 / (ms / R * 1e-3) / 1e9);
        CK(cudaGetLastError());
    };
#define SET(k, bytes) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
    SET(k_cpasync<1>, 32 * STAGE); SET(k_cpasync<2>, 64 * STAGE); SET(k_tma<1>, 32 * STAGE); SET(k_tma<2>, 64 * STAGE);
    SET(k_cpasync<3>, 72 * STAGE); SET(k_tma<3>, 72 * STAGE);
    for (int warps : {32, 16}) {
        char nm[128];
        snprintf(nm, 128, "cp.async 1 stage, %d warps/SM", warps);
        run(nm, [&] { k_cpasync<1><<<148, 1024, warps * STAGE>>>(val, idx, nchunks, warps, out); });
        snprintf(nm, 128, "cp.async 2 stages, %d warps/SM", warps);
        run(nm, [&] { k_cpasync<2><<<148, 1024, warps * 2 * STAGE>>>(val, idx, nchunks, warps, out); });
        snprintf(nm, 128, "TMA bulk 1 stage, %d warps/SM", warps);
        run(nm, [&] { k_tma<1><<<148, 1024, warps * STAGE>>>(val, idx, nchunks, warps, out); });
        snprintf(nm, 128, "TMA bulk 2 stages, %d warps/SM", warps);
        run(nm, [&] { k_tma<2><<<148, 1024, warps * 2 * STAGE>>>(val, idx, nchunks, warps, out); });
        snprintf(nm, 128, "LDG.128 to registers, %d warps/SM", warps);
        run(nm, [&] { k_ldg<<<148, 1024>>>(val, idx, nchunks, warps, out); });
    }
    run("cp.async 3 stages, 24 warps/SM", [&] { k_cpasync<3><<<148, 1024, 72 * STAGE>>>(val, idx, nchunks, 24, out); });
    run("TMA bulk 3 stages, 24 warps/SM", [&] { k_tma<3><<<148, 1024, 72 * STAGE>>>(val, idx, nchunks, 24, out); });
    return 0;
}
