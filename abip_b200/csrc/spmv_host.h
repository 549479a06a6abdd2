// spmv_host.h -- host-side helpers shared by the LP and QCP engines: CUDA error macro, SpMV plan builder,
// padded uploads, CSR transposition.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "lp_device.cuh"

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t _e = (call);                                                                         \
        if (_e != cudaSuccess) {                                                                         \
            fprintf(stderr, "[abip_gpu] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, \
                    __LINE__, cudaGetErrorString(_e));                                                   \
            return -1;                                                                                   \
        }                                                                                                \
    } while (0)

static inline int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Row-length statistics -> SpMV plan for a persistent grid of W warps (see Csr in lp_device.cuh):
// contiguous, cost-balanced row ranges per warp (cost = nnz + 2 per row), cut into chunks of <= kChunk nonzeros and
// <= kChunk rows; rows longer than kChunk become single-row chunks (warp-per-row 128-bit path).  The lane count of
// the shared-memory row reduction follows the mean row length: L ~ 16 * mean / kChunk rounded up to a power of two
// (L = 1 for short rows, e.g. A' of an LP with ~5 nnz/column; L = 2 at ~25 nnz/row).
struct SpmvPlan {
    std::vector<int> warp_chunk;
    std::vector<int4> chunk;
    int lanes_log2 = 0;
    double mean = 0;
    int n_long = 0;
    int max_len = 0;
};

static inline void build_spmv_plan(const std::vector<int>& ptr, int nrows, int W, const char* env_lanes, SpmvPlan* P) {
    const long nnz = ptr[nrows];
    const double total_cost = (double)nnz + 2.0 * nrows;
    P->warp_chunk.assign(W + 1, 0);
    P->chunk.clear();
    P->n_long = 0;
    P->max_len = 0;
    int r = 0;
    for (int w = 0; w < W; ++w) {
        const double target = total_cost * (double)(w + 1) / (double)W;
        const int ra = r;
        while (r < nrows && ((double)ptr[r + 1] + 2.0 * (r + 1) <= target || w == W - 1)) ++r;
        int q = ra;
        while (q < r) {  // cut [ra, r) into chunks
            int q1 = q;
            int n = 0;
            while (q1 < r && (q1 - q) < kChunk) {
                const int len = ptr[q1 + 1] - ptr[q1];
                if (n + len > kChunk) break;
                n += len;
                ++q1;
            }
            if (q1 == q) {  // a row longer than kChunk: consecutive pieces, all in this warp
                const int len = ptr[q + 1] - ptr[q];
                for (int off = 0; off < len; off += kChunk) {
                    const int cnt = std::min(kChunk, len - off);
                    P->chunk.push_back(make_int4(q, ptr[q] + off, (off + cnt == len) ? -1 : 0, cnt));
                }
                P->n_long++;
                q = q + 1;
                continue;
            }
            P->chunk.push_back(make_int4(q, ptr[q], q1 - q, n));
            q = q1;
        }
        P->warp_chunk[w + 1] = (int)P->chunk.size();
    }
    for (int i = 0; i < nrows; ++i) P->max_len = std::max(P->max_len, ptr[i + 1] - ptr[i]);
    P->mean = nrows ? (double)nnz / nrows : 1.0;
    // measured at cfg2: fewer lanes per row win (each extra pass over the chunk costs more than a longer serial sum)
    const double want = 16.0 * P->mean / kChunk;
    int lg = 0;
    while (lg < 5 && (1 << lg) < want) ++lg;
    const int forced = env_int(env_lanes, -1);
    if (forced >= 0 && forced <= 5) lg = forced;
    P->lanes_log2 = lg;
}


// device copy of a CSR matrix + its plan
struct DevCsr {
    int *ptr = nullptr, *idx = nullptr, *wc = nullptr;
    int4* chunk = nullptr;
    double* val = nullptr;
    SpmvPlan plan;
    int nrows = 0;
    long nnz = 0;
    Csr view() const { return Csr{ptr, idx, val, nrows, wc, chunk, plan.lanes_log2}; }
    void release() {
        cudaFree(ptr); cudaFree(idx); cudaFree(wc); cudaFree(chunk); cudaFree(val);
        ptr = idx = wc = nullptr; chunk = nullptr; val = nullptr;
    }
};

template <class T>
static inline int upload_padded(T** dst, const std::vector<T>& src, cudaStream_t stream) {
    const size_t bytes = (src.size() + 8) * sizeof(T);  // +8: 16-byte aligned staging windows may over-read
    CK(cudaMalloc((void**)dst, bytes));
    CK(cudaMemsetAsync(*dst, 0, bytes, stream));
    if (!src.empty()) CK(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    return 0;
}

static inline int upload_csr(DevCsr* d, const std::vector<int>& ptr, const std::vector<int>& idx,
                             const std::vector<double>& val, int nrows, int W, const char* env_lanes,
                             cudaStream_t stream) {
    d->nrows = nrows;
    d->nnz = ptr[nrows];
    build_spmv_plan(ptr, nrows, W, env_lanes, &d->plan);
    if (upload_padded(&d->ptr, ptr, stream) || upload_padded(&d->idx, idx, stream) ||
        upload_padded(&d->val, val, stream) || upload_padded(&d->wc, d->plan.warp_chunk, stream) ||
        upload_padded(&d->chunk, d->plan.chunk, stream))
        return -1;
    CK(cudaStreamSynchronize(stream));  // the host vectors may go out of scope
    return 0;
}

// CSC (long indices, as passed through the reference ABI) -> CSR arrays of the same matrix (counting sort)
template <class IntT>
static inline int csc_to_csr(long nrows, long ncols, const IntT* Ap, const IntT* Ai, const double* Ax,
                             std::vector<int>* rptr, std::vector<int>* ridx, std::vector<double>* rval) {
    const long nnz = (long)Ap[ncols];
    rptr->assign(nrows + 1, 0);
    ridx->resize(nnz);
    rval->resize(nnz);
    for (long k = 0; k < nnz; ++k) {
        if (Ai[k] < 0 || Ai[k] >= nrows) return -1;
        (*rptr)[Ai[k] + 1]++;
    }
    for (long i = 0; i < nrows; ++i) (*rptr)[i + 1] += (*rptr)[i];
    std::vector<int> fill(rptr->begin(), rptr->end() - 1);
    for (long j = 0; j < ncols; ++j)
        for (long k = (long)Ap[j]; k < (long)Ap[j + 1]; ++k) {
            const int q = fill[Ai[k]]++;
            (*ridx)[q] = (int)j;
            (*rval)[q] = Ax[k];
        }
    return 0;
}
