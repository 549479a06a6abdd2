// order_host.h -- locality ordering of the rows and columns of A, decided at plan time (a pure function of the sparsity
// structure => bit-reproducible).
//
// Why (profiles/r02_spmv.md): on B200 the FP64 gathers x[col] of a CSR SpMV are bound by the number of
// distinct 128-byte lines one 32-lane load touches and by the 32-byte sectors they pull over the L2 -> SM path (5 M random
// gathers alone take as long as a whole SpMV pass), not by HBM.  locality_order() makes structurally identical
// ("shift-copy") rows / columns neighbours: block-replicated LPs (multi-commodity, multi-period, scenario trees) then put
// the copies of one row on neighbouring lanes and the copies of one column at consecutive addresses, so a gather touches a
// few lines instead of 32 (cfg2: 23.9 -> 11.9 sectors per request).  Matrices without replicated structure keep their order.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <thread>
#include <vector>

// fn(begin, end, thread id) over [0, n) on up to `threads` host threads (set-up of large problems only)
template <class Fn>
static inline void parallel_for(long n, int threads, Fn fn) {
    threads = (int)std::max<long>(1, std::min<long>(threads, n / 4096));
    if (threads <= 1) {
        fn(0L, n, 0);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        const long b = n * t / threads, e2 = n * (t + 1) / threads;
        pool.emplace_back([=] { fn(b, e2, t); });
    }
    for (auto& th : pool) th.join();
}

namespace sjds {  // (name kept from the sliced-JDS experiments of round 2, tools/ubench)

constexpr int kLongRow = 96;   // rows / columns longer than this are "dense": they take no part in the ordering keys

static inline uint64_t mix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// stable order of 0..n-1 by (group representative = smallest member index of the element's key class, index).
// O(n): the representative of a class is its first occurrence (open-addressing table key -> first index), the final
// order is a counting sort by representative.
static inline void order_by_class(const std::vector<uint64_t>& key, std::vector<int>* new2old) {
    const int n = (int)key.size();
    size_t cap = 16;
    while (cap < (size_t)n * 2) cap <<= 1;
    std::vector<int> slot(cap, -1);
    std::vector<int> rep(n), cnt(n + 1, 0);
    for (int i = 0; i < n; ++i) {
        size_t h = (size_t)(mix64(key[i]) & (cap - 1));
        for (;;) {
            const int j = slot[h];
            if (j < 0) { slot[h] = i; rep[i] = i; break; }
            if (key[j] == key[i]) { rep[i] = j; break; }
            h = (h + 1) & (cap - 1);
        }
        cnt[rep[i] + 1]++;
    }
    for (int i = 0; i < n; ++i) cnt[i + 1] += cnt[i];
    new2old->resize(n);
    for (int i = 0; i < n; ++i) (*new2old)[cnt[rep[i]]++] = i;
}

// Locality ordering (see file header).  Inputs: CSR(A) (m rows) and CSR(A') (n rows), column indices ascending.
//   a. row signature = hash(length, column differences to the first column): shift-copies of a row share it;
//   b. column key = multiset of the signatures of its (non-dense) rows; columns with equal keys are copies of each
//      other: classes are laid out in the order of their first member, members in index order;
//   c. row key = multiset of the classes of its (non-dense) columns; same layout rule.
// Dense rows / columns (longer than dense_thr) take no part in the keys and keep their relative place.
template <class ParFor>
static inline void locality_order(int m, int n, const std::vector<int>& a_ptr, const std::vector<int>& a_idx,
                                  const std::vector<int>& at_ptr, const std::vector<int>& at_idx, int dense_thr,
                                  std::vector<int>* row_new2old, std::vector<int>* col_new2old, ParFor par) {
    std::vector<uint64_t> rsig(m), ckey(n), rkey(m);
    par(m, [&](long r0, long r1, int) {
    for (long r = r0; r < r1; ++r) {
        const int a = a_ptr[r], b = a_ptr[r + 1];
        if (b - a > dense_thr || b == a) { rsig[r] = 0; continue; }
        uint64_t h = mix64((uint64_t)(b - a));
        for (int k = a + 1; k < b; ++k) h = mix64(h ^ (uint64_t)(uint32_t)(a_idx[k] - a_idx[a]));
        rsig[r] = h | 1ull;
    }
    });
    par(n, [&](long c0, long c1, int) {
    for (long c = c0; c < c1; ++c) {
        const int a = at_ptr[c], b = at_ptr[c + 1];
        if (b - a > dense_thr) { ckey[c] = mix64(0xc0ffeeull + (uint64_t)c); continue; }
        uint64_t h = 0;
        int cnt = 0;
        for (int k = a; k < b; ++k) {
            const uint64_t s = rsig[at_idx[k]];
            if (s) { h += mix64(s); ++cnt; }
        }
        ckey[c] = mix64(h ^ ((uint64_t)cnt << 56));
    }
    });
    order_by_class(ckey, col_new2old);
    // class id of a column = its key (collisions only merge classes, which is harmless)
    par(m, [&](long r0, long r1, int) {
    for (long r = r0; r < r1; ++r) {
        const int a = a_ptr[r], b = a_ptr[r + 1];
        if (b - a > dense_thr) { rkey[r] = mix64(0xabcdefull + (uint64_t)r); continue; }
        uint64_t h = 0;
        for (int k = a; k < b; ++k) {
            const int c = a_idx[k];
            if (at_ptr[c + 1] - at_ptr[c] <= dense_thr) h += mix64(ckey[c]);
        }
        rkey[r] = mix64(h ^ ((uint64_t)(b - a) << 56));
    }
    });
    order_by_class(rkey, row_new2old);
}


}  // namespace sjds
