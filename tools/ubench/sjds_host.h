// sjds_host.h -- EXPERIMENT (round 2, not part of the product): sliced-JDS layout + plan for tools/ubench/sjds_bench.cu.
// See profiles/r02_spmv.md for what was measured and why the engine keeps its CSR-stream kernel.
#pragma once
#include "order_host.h"

namespace sjds {
#ifndef ABIP_CH
#define ABIP_CH 256
#endif
constexpr int kPiece = ABIP_CH;  // nonzeros per piece of a long row (one stage of the device ring)
constexpr int kSkipLane = 0xff;  // meta high byte: this lane of the slice has no row (padding or long row)

// B = P_r A P_c in CSR: row i of B is row row_new2old[i] of A, column indices mapped through col_old2new; entries of a
// row keep the order of A (its summation order does not change).
static inline void permute_csr(int nrows, const std::vector<int>& ptr, const std::vector<int>& idx, const std::vector<double>& val,
                               const std::vector<int>& row_new2old, const std::vector<int>& col_old2new,
                               std::vector<int>* optr, std::vector<int>* oidx, std::vector<double>* oval) {
    optr->assign(nrows + 1, 0);
    oidx->resize(idx.size());
    oval->resize(val.size());
    for (int i = 0; i < nrows; ++i) {
        const int r = row_new2old[i];
        (*optr)[i + 1] = (*optr)[i] + (ptr[r + 1] - ptr[r]);
    }
    for (int i = 0; i < nrows; ++i) {
        const int r = row_new2old[i];
        int q = (*optr)[i];
        for (int k = ptr[r]; k < ptr[r + 1]; ++k, ++q) {
            (*oidx)[q] = col_old2new[idx[k]];
            if (!val.empty()) (*oval)[q] = val[k];
        }
    }
}

// One staging unit of the device kernel: a run of whole steps of ONE slice (or one piece of a long row) of at most
// kChunkElems elements; {first element, #elements, first row of the slice | piece slot, info}.
//   info: bits 0-7 #steps, bits 8-15 first step, bit 16 last chunk of its slice, bit 18 piece of a long row
constexpr int kChunkElems = ABIP_CH;
constexpr int kInfoLast = 1 << 16, kInfoLong = 1 << 18;
struct Chunk { int start, count, row0, info; };
struct LongRow { int row, slot0, npieces, pad; };

// Host image of one matrix in SJDS form + its work plan for a persistent grid of G CTAs x wpc warps.
struct Host {
    int nrows = 0, ncols = 0, nslices = 0;
    long nnz = 0;
    std::vector<double> val;            // slices (step-major), then the long rows (row-major)
    std::vector<int> idx;
    std::vector<int> sptr;              // [nslices + 1], multiples of 4 (host only)
    std::vector<unsigned short> meta;   // [32 * nslices]: length | source lane << 8 (kSkipLane: nothing to do)
    std::vector<int> src_of;            // [stored elements] position of every element in the CSR arrays (-1: padding)
    // plan
    std::vector<Chunk> chunk;           // grouped by warp of the persistent grid
    std::vector<int> warp_chunk;        // [G * wpc + 1]
    std::vector<LongRow> long_rows;     // grouped by CTA
    std::vector<int> cta_long;          // [G + 1]
    int n_pieces = 0, n_long = 0, max_len = 0;
    double mean = 0;
};

static inline void build(int nrows, int ncols, const std::vector<int>& ptr, const std::vector<int>& idx,
                         const std::vector<double>& val, int G, int wpc, Host* H) {
    H->nrows = nrows;
    H->ncols = ncols;
    H->nnz = ptr[nrows];
    H->nslices = (nrows + 31) / 32;
    H->sptr.assign(H->nslices + 1, 0);
    H->meta.assign((size_t)32 * H->nslices, (unsigned short)(kSkipLane << 8));
    H->val.clear();
    H->idx.clear();
    H->src_of.clear();
    H->val.reserve(H->nnz + 4 * (size_t)H->nslices);
    H->idx.reserve(H->nnz + 4 * (size_t)H->nslices);
    H->src_of.reserve(H->nnz + 4 * (size_t)H->nslices);
    std::vector<std::vector<Chunk>> slice_chunks(H->nslices);
    std::vector<long> slice_cost(H->nslices, 0);
    std::vector<int> longs;
    H->max_len = 0;
    for (int s = 0; s < H->nslices; ++s) {
        int lanes[32], len[32], nl = 0;
        for (int q = 0; q < 32; ++q) {
            const int r = 32 * s + q;
            if (r >= nrows) break;
            const int L = ptr[r + 1] - ptr[r];
            H->max_len = std::max(H->max_len, L);
            if (L > kLongRow) { longs.push_back(r); continue; }
            lanes[nl] = q;
            len[nl] = L;
            ++nl;
        }
        int ord[32];
        for (int q = 0; q < nl; ++q) ord[q] = q;
        std::stable_sort(ord, ord + nl, [&](int a, int b) { return len[a] > len[b]; });
        for (int q = 0; q < nl; ++q)
            H->meta[(size_t)32 * s + q] = (unsigned short)(len[ord[q]] | (lanes[ord[q]] << 8));
        const int maxlen = nl ? len[ord[0]] : 0;
        // steps -> chunks of whole steps (a chunk's copy window starts at the previous multiple of 4 elements)
        Chunk cur{(int)H->val.size(), 0, 32 * s, 0};
        int cur_steps = 0, cur_j0 = 0;
        auto flush = [&](bool last) {
            cur.info = cur_steps | (cur_j0 << 8) | (last ? kInfoLast : 0);
            slice_chunks[s].push_back(cur);
            slice_cost[s] += cur.count + 8 * cur_steps + 24;
        };
        for (int j = 0; j < maxlen; ++j) {
            int cnt = 0;
            while (cnt < nl && len[ord[cnt]] > j) ++cnt;
            if (cur_steps > 0 && ((cur.start & 3) + cur.count + cnt > kChunkElems || cur_steps == 255)) {
                flush(false);
                cur = Chunk{(int)H->val.size(), 0, 32 * s, 0};
                cur_steps = 0;
                cur_j0 = j;
            }
            for (int q = 0; q < cnt; ++q) {
                const int r = 32 * s + lanes[ord[q]];
                const int k = ptr[r] + j;
                H->val.push_back(val.empty() ? 0.0 : val[k]);
                H->idx.push_back(idx[k]);
                H->src_of.push_back(k);
            }
            cur.count += cnt;
            ++cur_steps;
        }
        flush(true);  // (a slice without nonzeros still gets one chunk: its rows need their epilogue)
        while (H->val.size() & 3) { H->val.push_back(0.0); H->idx.push_back(0); H->src_of.push_back(-1); }
        H->sptr[s + 1] = (int)H->val.size();
    }
    // long rows: row-major behind the slices; pieces of <= kPiece nonzeros, dealt with their row to one CTA
    H->n_long = (int)longs.size();
    std::vector<std::vector<Chunk>> cta_pieces(G);
    std::vector<long> cta_cost(G, 0);
    H->long_rows.clear();
    H->cta_long.assign(G + 1, 0);
    {
        std::vector<std::vector<int>> cta_rows(G);
        for (size_t q = 0; q < longs.size(); ++q) cta_rows[q % G].push_back(longs[q]);
        int slot = 0;
        for (int b = 0; b < G; ++b) {
            for (int r : cta_rows[b]) {
                const int L = ptr[r + 1] - ptr[r];
                const int np = (L + kPiece - 1) / kPiece;
                const int per = (((L + np - 1) / np) + 3) & ~3;
                H->long_rows.push_back(LongRow{r, slot, np, 0});
                const int base = (int)H->val.size();
                for (int k = ptr[r]; k < ptr[r + 1]; ++k) {
                    H->val.push_back(val.empty() ? 0.0 : val[k]);
                    H->idx.push_back(idx[k]);
                    H->src_of.push_back(k);
                }
                while (H->val.size() & 3) { H->val.push_back(0.0); H->idx.push_back(0); H->src_of.push_back(-1); }
                for (int i = 0, off = 0; i < np; ++i, off += per) {
                    const int cnt = std::min(per, L - off);
                    cta_pieces[b].push_back(Chunk{base + off, cnt, slot + i, kInfoLong});
                    cta_cost[b] += cnt + 8 * ((cnt + 31) / 32) + 24;
                }
                slot += np;
            }
            H->cta_long[b + 1] = (int)H->long_rows.size();
        }
        H->n_pieces = slot;
    }
    // contiguous slice ranges so that every CTA carries the same total cost (its long rows included); inside a CTA the
    // units (pieces first, then slices) are dealt round-robin to the warps, each warp's chunks stored consecutively
    long total = 0;
    for (int s = 0; s < H->nslices; ++s) total += slice_cost[s];
    for (int b = 0; b < G; ++b) total += cta_cost[b];
    H->chunk.clear();
    H->warp_chunk.assign((size_t)G * wpc + 1, 0);
    {
        int s = 0;
        long acc = 0;
        std::vector<std::vector<Chunk>> per_warp(wpc);
        for (int b = 0; b < G; ++b) {
            acc += cta_cost[b];
            const int s_begin = s;
            const double target = (double)total * (b + 1) / G;
            while (s < H->nslices && (b == G - 1 || (double)(acc + slice_cost[s] / 2) <= target)) acc += slice_cost[s++];
            for (auto& v : per_warp) v.clear();
            int u = 0;
            for (const Chunk& c : cta_pieces[b]) per_warp[u++ % wpc].push_back(c);
            for (int q = s_begin; q < s; ++q, ++u)
                for (const Chunk& c : slice_chunks[q]) per_warp[u % wpc].push_back(c);
            for (int w = 0; w < wpc; ++w) {
                for (const Chunk& c : per_warp[w]) H->chunk.push_back(c);
                H->warp_chunk[(size_t)b * wpc + w + 1] = (int)H->chunk.size();
            }
        }
    }
    H->mean = nrows ? (double)H->nnz / nrows : 0.0;
}

// CPU model of the gather traffic of one pass: distinct 128-byte lines and 32-byte sectors per 32-lane gather request
struct GatherStats { double requests = 0, lines = 0, sectors = 0, lanes = 0; };
static inline GatherStats gather_stats(const Host& H) {
    GatherStats g;
    for (int s = 0; s < H.nslices; ++s) {
        int len[32];
        for (int q = 0; q < 32; ++q) len[q] = H.meta[(size_t)32 * s + q] & 0xff;
        int base = H.sptr[s];
        for (int j = 0; j < len[0]; ++j) {
            int cnt = 0;
            while (cnt < 32 && len[cnt] > j) ++cnt;
            long ln[32], sc[32];
            for (int q = 0; q < cnt; ++q) { ln[q] = H.idx[base + q] >> 4; sc[q] = H.idx[base + q] >> 2; }
            std::sort(ln, ln + cnt);
            std::sort(sc, sc + cnt);
            g.requests += 1;
            g.lanes += cnt;
            g.lines += std::unique(ln, ln + cnt) - ln;
            g.sectors += std::unique(sc, sc + cnt) - sc;
            base += cnt;
        }
    }
    for (const Chunk& p : H.chunk) {
        if (!(p.info & kInfoLong)) continue;
        for (int o = 0; o < p.count; o += 32) {
            const int cnt = std::min(32, p.count - o);
            long ln[32], sc[32];
            for (int q = 0; q < cnt; ++q) { ln[q] = H.idx[p.start + o + q] >> 4; sc[q] = H.idx[p.start + o + q] >> 2; }
            std::sort(ln, ln + cnt);
            std::sort(sc, sc + cnt);
            g.requests += 1;
            g.lanes += cnt;
            g.lines += std::unique(ln, ln + cnt) - ln;
            g.sectors += std::unique(sc, sc + cnt) - sc;
        }
    }
    return g;
}

}  // namespace sjds
