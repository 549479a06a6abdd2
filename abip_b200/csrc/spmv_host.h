// spmv_host.h -- host-side helpers shared by the LP and QCP engines: CUDA error macro, SpMV plan builder,
// padded uploads, CSR transposition.
#pragma once
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <cuda_runtime.h>
#include "lp_device.cuh"
#include "order_host.h"

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
    std::vector<int> cta_row;     // [G+1] contiguous row range of each CTA (deal == 0)
    std::vector<double> row_cost; // model cost of every row (kept for the measured balance, tune_balance() in lp_engine.cu)
    int n_pieces = 0;
    int lanes_log2 = 0;
    double mean = 0;
    int n_long = 0;
    int max_len = 0;
};

// Cost of a row for the balance of the CTAs: nnz + 2 (stream + row overhead) + kLineCost per 128-byte line of the gathered
// vector that the row is the first of its neighbourhood (~ one chunk) to touch.  A gather costs L1 tag look-ups and L2
// sectors per distinct line, not per element: after the locality ordering a chunk of 20 structurally identical rows
// touches ~25 lines, a piece of a long random row ~220 -- measured 16.5 vs 34 us per pass for CTAs holding the same
// number of nonzeros; the weight was then tuned on the whole solve (profiles/r02_spmv.md: 0.5-0.7 best, > 1 worse).
constexpr double kLineCost = 0.5;
static inline void row_line_costs(const std::vector<int>& ptr, const int* idx, int nrows, std::vector<double>* cost) {
    cost->resize(nrows);
    constexpr int kSet = 2048;  // open addressing, epoch-tagged: reset = epoch increment
    std::vector<int> key(kSet, -1), tag(kSet, -1);
    int epoch = 0, window = 0;
    const char* wenv = getenv("ABIP_GPU_LINE_COST");  // tuning knob
    const double w = (wenv && *wenv) ? atof(wenv) : kLineCost;
    for (int r = 0; r < nrows; ++r) {
        const int a = ptr[r], b = ptr[r + 1];
        int fresh = 0;
        if (b - a > kChunk) {
            fresh = b - a;  // long rows: every piece is a chunk of its own, assume no reuse
        } else {
            for (int k = a; k < b; ++k) {
                const int line = idx[k] >> 4;
                unsigned h = ((unsigned)line * 2654435761u) >> 21;  // 11 bits
                for (;;) {
                    if (tag[h] != epoch) { tag[h] = epoch; key[h] = line; ++fresh; break; }
                    if (key[h] == line) break;
                    h = (h + 1) & (kSet - 1);
                }
            }
            window += b - a;
            if (window >= kChunk) { window = 0; ++epoch; }
        }
        (*cost)[r] = (double)(b - a) + 2.0 + w * fresh;
    }
}

// cost_override: per-row costs that replace the model (measured balance: the model costs rescaled CTA by CTA with the
// measured pass times); keep_cost: leave the model costs in P->row_cost.
static inline void build_spmv_plan(const std::vector<int>& ptr, int nrows, int W, const char* env_lanes, SpmvPlan* P,
                                   const int* idx = nullptr, const std::vector<double>* cost_override = nullptr,
                                   bool keep_cost = false) {
    const long nnz = ptr[nrows];
    const int G = std::max(1, W / kWarps);
    // prefix of the row costs (contiguous ranges, deal == 0)
    std::vector<double> cum(nrows + 1, 0.0);
    if (cost_override) {
        for (int r = 0; r < nrows; ++r) cum[r + 1] = cum[r] + (*cost_override)[r];
    } else if (idx && G > 1) {
        std::vector<double> rc;
        row_line_costs(ptr, idx, nrows, &rc);
        for (int r = 0; r < nrows; ++r) cum[r + 1] = cum[r] + rc[r];
        if (keep_cost) P->row_cost.swap(rc);
    } else {
        for (int r = 0; r < nrows; ++r) cum[r + 1] = (double)ptr[r + 1] + 2.0 * (r + 1);
    }
    const double total_cost = cum[nrows];
    P->warp_chunk.assign(W + 1, 0);
    P->chunk.clear();
    P->cta_long.assign(G + 1, 0);
    P->cta_row.assign(G + 1, 0);
    P->long_rows.clear();
    P->n_pieces = 0;
    P->n_long = 0;
    P->max_len = 0;
    // ABIP_GPU_PLAN_DEAL: how rows are assigned to the CTAs of the persistent grid.
    //   0 (default): contiguous range of rows per CTA, balanced by the locality-aware row costs above;
    //   2: long rows go to the least-loaded CTA, longest first (they have no locality); the chunks of the
    //      short rows are dealt in groups of kDealGroup consecutive chunks, each group to the least-loaded CTA -- every
    //      CTA gets the same mix of row kinds (phase time = slowest CTA; with contiguous ranges the CTAs that held the
    //      long rows took 2x the mean after the locality ordering), all CTAs walk through the matrix region by region,
    //      and neighbouring rows still share a CTA (L1 reuse of the gathered lines) -- measured slower at cfg2: the random
    //      gathers of the long rows then disturb the L1 reuse of the structured rows in EVERY CTA (mean busy time per
    //      pass 25.5 -> 33.6 us);
    //   1: single chunks dealt round-robin over the CTAs.
    // Measured at cfg2: profiles/r01_spmv_variants.md, profiles/r02_spmv.md.
    const int deal = env_int("ABIP_GPU_PLAN_DEAL", 0);
    std::vector<std::vector<int4>> cta_units(G);
    std::vector<std::vector<int4>> cta_lr(G);
    auto emit_long = [&](int owner, int q) {
        const int len0 = ptr[q + 1] - ptr[q];
        const int np = (len0 + kChunk - 1) / kChunk;
        const int per = (len0 + np - 1) / np;
        cta_lr[owner].push_back(make_int4(q, 0, np, 0));
        for (int i = 0, off = 0; i < np; ++i, off += per)
            cta_units[owner].push_back(make_int4(q, ptr[q] + off, -i - 1, std::min(per, len0 - off)));
        P->n_long++;
    };
    if (deal == 2) {
        constexpr int kDealGroup = 4;
        const long rounds = std::max(1L, (nnz + (long)W * kChunk - 1) / ((long)W * kChunk));
        const int want = (int)std::min<long>(kChunk, std::max<long>(32, (nnz + rounds * W - 1) / (rounds * W)));
        std::vector<double> load(G, 0.0);
        auto least = [&]() {
            int bb = 0;
            for (int b = 1; b < G; ++b)
                if (load[b] < load[bb]) bb = b;
            return bb;
        };
        std::vector<int> longs;
        for (int q = 0; q < nrows; ++q)
            if (ptr[q + 1] - ptr[q] > kChunk) longs.push_back(q);
        std::stable_sort(longs.begin(), longs.end(), [&](int x, int y) { return ptr[x + 1] - ptr[x] > ptr[y + 1] - ptr[y]; });
        for (int q : longs) {
            const int b = least();
            emit_long(b, q);
            load[b] += 1.25 * (ptr[q + 1] - ptr[q]) + 64.0;
        }
        std::vector<int4> grp;
        double grp_cost = 0;
        auto flush = [&]() {
            if (grp.empty()) return;
            const int b = least();
            for (const int4& u : grp) cta_units[b].push_back(u);
            load[b] += grp_cost;
            grp.clear();
            grp_cost = 0;
        };
        int q = 0;
        while (q < nrows) {
            if (ptr[q + 1] - ptr[q] > kChunk) { ++q; continue; }
            int q1 = q, n = 0;
            while (q1 < nrows && (q1 - q) < kChunkRows && n < want) {
                const int len = ptr[q1 + 1] - ptr[q1];
                if (len > kChunk || n + len > kChunk) break;
                n += len;
                ++q1;
            }
            grp.push_back(make_int4(q, ptr[q], q1 - q, n));
            grp_cost += n + 2.0 * (q1 - q) + 48.0;
            if ((int)grp.size() == kDealGroup) flush();
            q = q1;
        }
        flush();
    } else {
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
                while (r < nrows && (cum[r + 1] <= target || b == G - 1)) ++r;
                P->cta_row[b + 1] = r;
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
                    emit_long(owner, q);
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


// device copy of a CSR matrix + its plan
struct DevCsr {
    int *ptr = nullptr, *idx = nullptr, *wc = nullptr, *cta_long = nullptr;
    int4 *chunk = nullptr, *long_rows = nullptr;
    double *val = nullptr, *long_part = nullptr;
    SpmvPlan plan;
    int nrows = 0;
    long nnz = 0;
    Csr view() const {
        return Csr{ptr, idx, val, nrows, wc, chunk, plan.lanes_log2, cta_long, long_rows, long_part, 0, nullptr, nullptr, nullptr};
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
