// Host check of the chunk-local JDS layout (abip_b200/csrc/order_host.h: jds_sort_chunks) against the addressing of the
// lane-per-row SpMV path (abip_b200/csrc/lp_device.cuh, Csr::jds): a warp of 32 lanes is emulated step by step with the
// same position arithmetic (base of step j = number of entries of the steps before it, lane l -> rows l and l + 32), and
// the row sums are compared with a plain CSR product in the original row order.  Exit code 0 = identical.
#include "../../abip_b200/csrc/order_host.h"
#include <cstdio>
#include <cstdlib>
#include <random>

struct D4 { int x, y, z, w; };

int main(int argc, char** argv) {
    const int nrows = argc > 1 ? atoi(argv[1]) : 5000;
    const int maxlen = argc > 2 ? atoi(argv[2]) : 12;
    const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
    const int ncols = 3000, kChunk = 252, kRows = 64;
    std::mt19937 rng(seed);
    std::vector<int> ptr(nrows + 1, 0), idx, src, n2o(nrows);
    std::vector<double> val;
    for (int r = 0; r < nrows; ++r) {
        const int len = (int)(rng() % (unsigned)(maxlen + 1));  // empty rows included
        for (int k = 0; k < len; ++k) {
            idx.push_back((int)(rng() % ncols));
            val.push_back((double)(rng() % 2001) / 1000.0 - 1.0);
        }
        ptr[r + 1] = (int)idx.size();
        n2o[r] = r;
    }
    const int nnz = (int)idx.size();
    src.resize(nnz);
    for (int k = 0; k < nnz; ++k) src[k] = k;
    std::vector<double> x(ncols);
    for (double& v : x) v = (double)(rng() % 1000) / 500.0 - 1.0;
    std::vector<double> y_ref(nrows, 0.0);
    for (int r = 0; r < nrows; ++r) {
        double acc = 0.0;
        for (int k = ptr[r]; k < ptr[r + 1]; ++k) acc = acc + val[k] * x[idx[k]];
        y_ref[r] = acc;
    }
    // chunks of whole rows: <= kChunk nonzeros, <= kRows rows (random target sizes, like the plan's `want`)
    std::vector<D4> chunks;
    for (int q = 0; q < nrows;) {
        const int want = 32 + (int)(rng() % 221);
        int q1 = q, n = 0;
        while (q1 < nrows && q1 - q < kRows && n < want) {
            const int len = ptr[q1 + 1] - ptr[q1];
            if (n + len > kChunk) break;
            n += len;
            ++q1;
        }
        chunks.push_back(D4{q, ptr[q], q1 - q, n});
        q = q1;
    }
    const std::vector<int> ptr_before = ptr;
    auto par = [&](long cnt, auto fn) { parallel_for(cnt, 4, fn); };
    sjds::jds_sort_chunks(chunks, ptr, idx, src, n2o, par);
    if (ptr[0] != 0 || ptr[nrows] != nnz) { printf("pointer ends moved\n"); return 1; }
    // n2o must stay a permutation, chunk by chunk
    {
        std::vector<int> seen(nrows, 0);
        for (int r = 0; r < nrows; ++r) seen[n2o[r]]++;
        for (int r = 0; r < nrows; ++r) if (seen[r] != 1) { printf("n2o is not a permutation\n"); return 1; }
    }
    std::vector<double> y(nrows, 0.0);
    for (const D4& d : chunks) {
        const int row0 = d.x, s = d.y, nr = d.z;
        int len0[32], len1[32];
        double acc0[32], acc1[32];
        int mx = 0;
        for (int l = 0; l < 32; ++l) {
            len0[l] = l < nr ? ptr[row0 + l + 1] - ptr[row0 + l] : 0;
            len1[l] = l + 32 < nr ? ptr[row0 + l + 33] - ptr[row0 + l + 32] : 0;
            acc0[l] = acc1[l] = 0.0;
            mx = std::max(mx, std::max(len0[l], len1[l]));
        }
        for (int l = 0; l + 1 < nr; ++l)
            if (ptr[row0 + l + 2] - ptr[row0 + l + 1] > ptr[row0 + l + 1] - ptr[row0 + l]) { printf("rows not sorted\n"); return 1; }
        int base = 0;
        for (int j = 0; j < mx; ++j) {
            int c0 = 0, c1 = 0;
            for (int l = 0; l < 32; ++l) {
                c0 += j < len0[l];
                c1 += j < len1[l];
            }
            for (int l = 0; l < 32; ++l) {
                if (j < len0[l]) {
                    const int p = s + base + l;
                    if (p >= s + d.w) { printf("position outside the chunk\n"); return 1; }
                    acc0[l] = acc0[l] + val[src[p]] * x[idx[p]];
                }
                if (j < len1[l]) {
                    const int p = s + base + 32 + l;
                    if (p >= s + d.w) { printf("position outside the chunk\n"); return 1; }
                    acc1[l] = acc1[l] + val[src[p]] * x[idx[p]];
                }
            }
            base += c0 + c1;
        }
        if (base != d.w) { printf("step counts do not add up to the chunk size\n"); return 1; }
        for (int l = 0; l < 32; ++l) {
            if (l < nr) y[row0 + l] = acc0[l];
            if (l + 32 < nr) y[row0 + l + 32] = acc1[l];
        }
    }
    for (int r = 0; r < nrows; ++r)
        if (y[r] != y_ref[n2o[r]]) { printf("row %d (old %d): %.17g != %.17g\n", r, n2o[r], y[r], y_ref[n2o[r]]); return 1; }
    for (int r = 0; r < nrows; ++r)
        if (ptr[r + 1] - ptr[r] != ptr_before[n2o[r] + 1] - ptr_before[n2o[r]]) { printf("row length mismatch\n"); return 1; }
    printf("ok: %d rows, %d nonzeros, %zu chunks\n", nrows, nnz, chunks.size());
    return 0;
}
