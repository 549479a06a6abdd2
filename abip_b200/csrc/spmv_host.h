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

// Stream-ordered device allocations (cudaMallocAsync / cudaFreeAsync): cudaMalloc and cudaFree synchronise the whole
// device, which serialised the concurrent engines of a batch of small LPs.  The pool keeps freed memory.
static inline cudaError_t dev_alloc(void** p, size_t bytes, cudaStream_t s) {
    static thread_local int pool_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (pool_dev != dev) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        pool_dev = dev;
    }
    return cudaMallocAsync(p, bytes, s);
}
static inline void dev_free(void* p, cudaStream_t s) {
    if (p) cudaFreeAsync(p, s);
}

static inline int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// Row-length statistics -> SpMV plan for a persistent grid of W = G * kWarps warps (see Csr in lp_device.cuh):
//   1. every CTA owns a contiguous, cost-balanced range of rows (cost = nnz + 2 per row);
//   2. the range is cut into chunks of <= kChunk nonzeros and <= kChunkRows rows; the chunk size is chosen so that
//      the CTA holds a multiple of kWarps chunks of (nearly) equal size; a row longer than kChunk is cut into
//      equal pieces, each piece a chunk of its own (d.z = -(piece slot) - 1): the piece sums go to a scratch array
//      and are added in piece order by one thread after the CTA has finished its chunks;
//   3. the chunks of a CTA are dealt round-robin to its warps (measured: with contiguous per-warp ranges the slowest
//      warp of a CTA took 1.7x the mean), and the descriptor array is stored grouped by warp.
// Everything is a pure function of the row pointers => results are bit-reproducible run to run.
// The lane count of the shared-memory row reduction follows the mean row length: L ~ 16 * mean / kChunk rounded up
// to a power of two (L = 1 for short rows, e.g. A' of an LP with ~5 nnz/column; L = 2 at ~25 nnz/row).
struct SpmvPlan {
    std::vector<int> warp_chunk;  // [W+1]
    std::vector<int4> chunk;      // grouped by warp
    std::vector<int> cta_long;    // [G+1] range of long_rows per CTA
    std::vector<int4> long_rows;  // {row, first piece slot, #pieces, 0}
    int n_pieces = 0;
    int lanes_log2 = 0;
    double mean = 0;
    int n_long = 0;
    int max_len = 0;
};

static inline void build_spmv_plan(const std::vector<int>& ptr, int nrows, int W, const char* env_lanes, SpmvPlan* P) {
    const long nnz = ptr[nrows];
    const int G = std::max(1, W / kWarps);
    const double total_cost = (double)nnz + 2.0 * nrows;
    P->warp_chunk.assign(W + 1, 0);
    P->chunk.clear();
    P->cta_long.assign(G + 1, 0);
    P->long_rows.clear();
    P->n_pieces = 0;
    P->n_long = 0;
    P->max_len = 0;
    // ABIP_GPU_PLAN_DEAL=0 (default): contiguous, cost-balanced row range per CTA (keeps the L1 locality of
    // neighbouring rows); =1: chunks are dealt round-robin over the CTAs of the grid, so every CTA gets the same mix of
    // row kinds.  Measured at cfg2 (profiles/r01_spmv_variants.md): 156.4 vs 150.1 ADMM it/s.
    const bool deal = env_int("ABIP_GPU_PLAN_DEAL", 0) != 0;
    std::vector<std::vector<int4>> cta_units(G);
    std::vector<std::vector<int4>> cta_lr(G);
    {
        const long rounds = std::max(1L, (nnz + (long)W * kChunk - 1) / ((long)W * kChunk));
        const int want_all = (int)std::min<long>(kChunk, std::max<long>(32, (nnz + rounds * W - 1) / (rounds * W)));
        int r = 0;
        long group = 0;
        for (int b = 0; b < G; ++b) {
            int ra = r, want = want_all;
            if (deal) {
                if (b > 0) break;  // one pass over all rows
                r = nrows;
            } else {
                const double target = total_cost * (double)(b + 1) / (double)G;
                while (r < nrows && ((double)ptr[r + 1] + 2.0 * (r + 1) <= target || b == G - 1)) ++r;
                const long nnz_cta = (long)ptr[r] - ptr[ra];
                const long rc = std::max(1L, (nnz_cta + (long)kWarps * kChunk - 1) / ((long)kWarps * kChunk));
                want = (int)std::min<long>(kChunk, std::max<long>(32, (nnz_cta + rc * kWarps - 1) / (rc * kWarps)));
            }
            int q = ra;
            while (q < r) {
                const int owner = deal ? (int)(group % G) : b;
                ++group;
                const int len0 = ptr[q + 1] - ptr[q];
                if (len0 > kChunk) {  // long row: equal pieces, all in one CTA
                    const int np = (len0 + kChunk - 1) / kChunk;
                    const int per = (len0 + np - 1) / np;
                    cta_lr[owner].push_back(make_int4(q, 0, np, 0));
                    for (int i = 0, off = 0; i < np; ++i, off += per)
                        cta_units[owner].push_back(make_int4(q, ptr[q] + off, -i - 1, std::min(per, len0 - off)));
                    P->n_long++;
                    ++q;
                    continue;
                }
                int q1 = q, n = 0;
                while (q1 < r && (q1 - q) < kChunkRows && n < want) {
                    const int len = ptr[q1 + 1] - ptr[q1];
                    if (len > kChunk || n + len > kChunk) break;
                    n += len;
                    ++q1;
                }
                cta_units[owner].push_back(make_int4(q, ptr[q], q1 - q, n));
                q = q1;
            }
        }
    }
    for (int b = 0; b < G; ++b) {
        // piece slots: consecutive per long row, in the order of the CTA's long-row list
        std::vector<int4>& units = cta_units[b];
        {
            size_t li = 0;
            for (size_t u = 0; u < units.size(); ++u) {
                if (units[u].z >= 0) continue;
                if (units[u].z == -1) {  // first piece of the next long row of this CTA
                    cta_lr[b][li].y = P->n_pieces;
                    P->n_pieces += cta_lr[b][li].z;
                    ++li;
                }
                const int first = cta_lr[b][li - 1].y;
                units[u].z = -(first + (-units[u].z - 1)) - 1;
            }
        }
        for (const int4& lr : cta_lr[b]) P->long_rows.push_back(lr);
        P->cta_long[b + 1] = (int)P->long_rows.size();
        for (int w = 0; w < kWarps; ++w) {
            for (size_t u = w; u < units.size(); u += kWarps) P->chunk.push_back(units[u]);
            P->warp_chunk[(size_t)b * kWarps + w + 1] = (int)P->chunk.size();
        }
    }
    for (int i = 0; i < nrows; ++i) P->max_len = std::max(P->max_len, ptr[i + 1] - ptr[i]);
    P->mean = nrows ? (double)nnz / nrows : 1.0;
    // measured at cfg2: fewer lanes per row win (each extra pass over the chunk costs more than a longer serial sum)
    const double wantl = 16.0 * P->mean / kChunk;
    int lg = 0;
    while (lg < 5 && (1 << lg) < wantl) ++lg;
    const int forced = env_int(env_lanes, -1);
    if (forced >= 0 && forced <= 5) lg = forced;
    P->lanes_log2 = lg;
}


// Page cache plan (see kPageLog2 in lp_device.cuh): for every CTA of the persistent grid count the references of
// its slice of the matrix to each 256-byte page of the gathered vector, keep the `slots` most referenced pages with
// at least `min_refs` references (a page costs two coalesced wavefronts to load and saves about one wavefront per
// reference), and re-encode the column indices of the slice that fall into a kept page as kPcFlag | (slot * 32 +
// offset).  Deterministic: ties are broken by page id.
struct PageCache {
    std::vector<int> npages;  // [G]
    std::vector<int> pages;   // [G * stride]
    int stride = 0;
    long hits = 0;            // nonzeros whose gather goes to shared memory
};

static inline void build_page_cache(const SpmvPlan& P, std::vector<int>& idx, long ncols, int G, int slots, int min_refs,
                                    PageCache* out) {
    out->stride = slots;
    out->npages.assign(G, 0);
    out->pages.assign((size_t)G * slots, 0);
    out->hits = 0;
    if (slots <= 0) return;
    const long npg = (ncols + kPageDoubles - 1) >> kPageLog2;
    std::vector<int> cnt(npg, 0), slot_of(npg, -1), touched;
    std::vector<std::pair<int, int>> cand;  // (-count, page)
    for (int b = 0; b < G; ++b) {
        const int c0 = P.warp_chunk[(size_t)b * kWarps], c1 = P.warp_chunk[(size_t)(b + 1) * kWarps];
        if (c0 >= c1) continue;
        touched.clear();
        for (int c = c0; c < c1; ++c)
            for (long k = P.chunk[c].y, e = k + P.chunk[c].w; k < e; ++k) {
                const int pg = idx[k] >> kPageLog2;
                if (cnt[pg]++ == 0) touched.push_back(pg);
            }
        cand.clear();
        for (int pg : touched)
            if (cnt[pg] >= min_refs) cand.emplace_back(-cnt[pg], pg);
        if ((int)cand.size() > slots) {
            std::nth_element(cand.begin(), cand.begin() + slots, cand.end());
            cand.resize(slots);
        }
        std::sort(cand.begin(), cand.end(), [](const std::pair<int, int>& a, const std::pair<int, int>& b2) { return a.second < b2.second; });
        int* pl = out->pages.data() + (size_t)b * slots;
        for (size_t i = 0; i < cand.size(); ++i) {
            pl[i] = cand[i].second;
            slot_of[cand[i].second] = (int)i;
        }
        out->npages[b] = (int)cand.size();
        for (int c = c0; c < c1; ++c)
            for (long k = P.chunk[c].y, e = k + P.chunk[c].w; k < e; ++k) {
                const int col = idx[k], sl = slot_of[col >> kPageLog2];
                if (sl >= 0) {
                    idx[k] = (int)(kPcFlag | (unsigned)((sl << kPageLog2) | (col & (kPageDoubles - 1))));
                    out->hits++;
                }
            }
        for (int pg : touched) { cnt[pg] = 0; slot_of[pg] = -1; }
    }
}

// device copy of a CSR matrix + its plan
struct DevCsr {
    int *ptr = nullptr, *idx = nullptr, *wc = nullptr, *cta_long = nullptr;
    int4 *chunk = nullptr, *long_rows = nullptr;
    double *val = nullptr, *long_part = nullptr;
    SpmvPlan plan;
    int nrows = 0;
    long nnz = 0;
    Csr view() const {
        return Csr{ptr, idx, val, nrows, wc, chunk, plan.lanes_log2, 0, 0, nullptr, nullptr, cta_long, long_rows, long_part, 0};
    }
    void release(cudaStream_t s = nullptr) {
        dev_free(ptr, s); dev_free(idx, s); dev_free(wc, s); dev_free(chunk, s); dev_free(val, s);
        dev_free(cta_long, s); dev_free(long_rows, s); dev_free(long_part, s);
        ptr = idx = wc = cta_long = nullptr; chunk = long_rows = nullptr; val = long_part = nullptr;
    }
};

template <class T>
static inline int upload_padded(T** dst, const std::vector<T>& src, cudaStream_t stream) {
    const size_t bytes = (src.size() + kPad) * sizeof(T);  // fixed-size staging windows over-read behind the arrays
    CK(dev_alloc((void**)dst, bytes, stream));
    CK(cudaMemsetAsync(*dst, 0, bytes, stream));
    if (!src.empty()) CK(cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, stream));
    return 0;
}

// long-row tables of a plan (nullptr when the matrix has no long rows)
static inline int upload_long_rows(const SpmvPlan& P, int** cta_long, int4** long_rows, double** long_part, cudaStream_t stream) {
    *cta_long = nullptr;
    *long_rows = nullptr;
    *long_part = nullptr;
    if (P.n_long == 0) return 0;
    if (upload_padded(cta_long, P.cta_long, stream) || upload_padded(long_rows, P.long_rows, stream)) return -1;
    CK(dev_alloc((void**)long_part, sizeof(double) * (P.n_pieces + 8), stream));
    CK(cudaMemsetAsync(*long_part, 0, sizeof(double) * (P.n_pieces + 8), stream));
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
        upload_padded(&d->chunk, d->plan.chunk, stream) || upload_long_rows(d->plan, &d->cta_long, &d->long_rows, &d->long_part, stream))
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
