// Micro-benchmark for round 2: cost of the in-chunk row sums of the SpMV (shared-memory wavefronts / bank conflicts).
// 32 warps per SM, each warp owns a 256-element window of FP64 products in shared memory and reduces it to row sums with
//   mode 0: one lane per row, serial LDS.64 over the row (shipped path for short rows, without the gather)
//   mode 1: two lanes per row + shuffle (shipped path for rows of ~25 nonzeros)
//   mode 2: sliced-ELL layout inside the chunk: element j of row r at [j * R + r] (R = rows of the chunk), one lane per
//           row -> consecutive lanes read consecutive addresses (conflict-free)
// Row lengths: uniform `len` (5 or 25) or jittered (len +- len/2).  Prints SM cycles per chunk (all 32 warps busy).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e_), __LINE__); exit(1);} } while (0)

constexpr int kWin = 256, kMaxRows = 64;

template <int MODE>
__global__ void __launch_bounds__(1024, 1) k_rowsum(const int* rowptr_all, int nrows, int reps, double* out, long long* cyc) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    double* vs = sm + w * (kWin + kMaxRows + 8);
    int* rp = reinterpret_cast<int*>(vs + kWin);
    for (int i = lane; i < kWin; i += 32) vs[i] = 1.0 + i * 1e-3;
    for (int i = lane; i <= nrows; i += 32) rp[i] = rowptr_all[i];
    __syncthreads();
    double acc_total = 0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (MODE == 0) {
            for (int rr = lane; rr < nrows; rr += 32) {
                const int a = rp[rr], b = rp[rr + 1];
                double acc = 0;
                for (int k = a; k < b; ++k) acc += vs[k];
                acc_total += acc;
            }
        } else if (MODE == 1) {
            const int sl = lane & 1, sub = lane >> 1;
            for (int base = 0; base < nrows; base += 16) {
                const int rr = base + sub;
                double acc = 0;
                if (rr < nrows) {
                    const int a = rp[rr], b = rp[rr + 1];
                    for (int k = a + sl; k < b; k += 2) acc += vs[k];
                }
                acc += __shfl_down_sync(0xffffffffu, acc, 1, 2);
                acc_total += acc;
            }
        } else {
            for (int r0 = 0; r0 < nrows; r0 += 32) {
                const int rr = r0 + lane, R = min(32, nrows - r0);
                if (rr < nrows) {
                    const int len = rp[rr + 1] - rp[rr];
                    const double* base = vs + rp[r0];          // slice start; element j of row (r0 + lane) at base[j * R + lane]
                    double acc = 0;
                    for (int j = 0; j < len; ++j) acc += base[j * R + lane];
                    acc_total += acc;
                }
            }
        }
        __syncwarp();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc_total == 1.2345e-300) out[0] = acc_total;
}

int main() {
    int* d_ptr; double* out; long long* cyc;
    CK(cudaMalloc(&d_ptr, 4 * (kMaxRows + 1))); CK(cudaMalloc(&out, 8)); CK(cudaMalloc(&cyc, 148 * 8));
    const int reps = 2000;
    const size_t smem = 32 * (kWin + kMaxRows + 8) * sizeof(double);
    CK(cudaFuncSetAttribute(k_rowsum<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_rowsum<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_rowsum<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int len : {5, 25}) {
        for (int jitter : {0, 1}) {
            std::vector<int> ptr(1, 0);
            srand(7);
            while ((int)ptr.size() <= kMaxRows) {
                int l = jitter ? len - len / 2 + rand() % (len + 1) : len;
                if (ptr.back() + l > 252) break;
                ptr.push_back(ptr.back() + l);
            }
            const int nrows = (int)ptr.size() - 1;
            ptr.resize(kMaxRows + 1, ptr.back());
            CK(cudaMemcpy(d_ptr, ptr.data(), 4 * (kMaxRows + 1), cudaMemcpyHostToDevice));
            double res[3];
            for (int mode = 0; mode < 3; ++mode) {
                for (int it = 0; it < 2; ++it) {
                    if (mode == 0) k_rowsum<0><<<148, 1024, smem>>>(d_ptr, nrows, reps, out, cyc);
                    else if (mode == 1) k_rowsum<1><<<148, 1024, smem>>>(d_ptr, nrows, reps, out, cyc);
                    else k_rowsum<2><<<148, 1024, smem>>>(d_ptr, nrows, reps, out, cyc);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<long long> c(148);
                CK(cudaMemcpy(c.data(), cyc, 148 * 8, cudaMemcpyDeviceToHost));
                double avg = 0; for (auto v : c) avg += v; avg /= 148;
                res[mode] = avg / reps / 32.0;   // SM cycles per chunk with 32 warps sharing the SM
            }
            printf("row length %2d%s, %2d rows/chunk: lane-per-row %6.1f | 2 lanes/row %6.1f | sliced-ELL %6.1f  SM cycles per chunk\n",
                   len, jitter ? " (jittered)" : "           ", nrows, res[0], res[1], res[2]);
        }
    }
    return 0;
}
