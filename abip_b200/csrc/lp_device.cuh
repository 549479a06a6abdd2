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

// Geometry chosen by measurement on B200 at cfg2 (profiles/r01_spmv_variants.md): two 512-thread CTAs per SM (32
// warps per SM either way; the CTA-level barriers of the reductions are cheaper), 252-nonzero chunks in fixed
// 256-element windows, one TMA-fed buffer per warp.
#ifndef ABIP_BLOCK
#define ABIP_BLOCK 512
#endif
#ifndef ABIP_MIN_BLOCKS_PER_SM
#define ABIP_MIN_BLOCKS_PER_SM 2
#endif
constexpr int kBlock = ABIP_BLOCK;
constexpr int kWarps = kBlock / 32;
constexpr int kMaxRed = 24;  // max scalars reduced between two grid barriers

// Virtual grid: the persistent device code never reads blockIdx / gridDim directly.  An ordinary launch maps them
// 1:1; the batched launch (k_batch, lp_engine.cu: one CTA = one independent small LP) presents every CTA as block 0 of a
// one-block grid, so the same phases run unchanged and the grid barrier degenerates to a CTA barrier.
__shared__ int2 s_vgrid;  // {block id, number of blocks}
__device__ __forceinline__ int VB() { return s_vgrid.x; }
__device__ __forceinline__ int VG() { return s_vgrid.y; }
__device__ __forceinline__ void vgrid_init(bool batched) {
    if (threadIdx.x == 0) s_vgrid = batched ? make_int2(0, 1) : make_int2((int)blockIdx.x, (int)gridDim.x);
    __syncthreads();
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// CSR matrix + the SpMV plan built on the host from its row-length statistics (lp_engine.cu: build_spmv_plan()).
//   * every warp of the persistent grid owns a contiguous range of rows with (nearly) equal cost, cut into
//     "chunks" of <= kChunk nonzeros and <= kChunk rows; descriptor = {first row, first nnz, #rows, #nnz};
//   * a row longer than kChunk is cut into pieces of <= kChunk nonzeros, all owned by the same warp:
//     #rows = 0 marks "row continues", #rows = -1 its last piece.
// The device arrays are padded by 8 elements so that 16-byte aligned copy windows may over-read.
#ifndef ABIP_CHUNK
#define ABIP_CHUNK 252
#endif
#ifndef ABIP_CHUNK_ROWS
#define ABIP_CHUNK_ROWS 64
#endif
constexpr int kChunk = ABIP_CHUNK;
constexpr int kChunkRows = ABIP_CHUNK_ROWS;  // rows per chunk (bounds the row-pointer window)
static_assert(ABIP_CHUNK_ROWS <= 64, "the lane-per-row path takes two rows per lane");
#ifndef ABIP_L1U
#define ABIP_L1U 4
#endif
constexpr int kL1U = ABIP_L1U;  // gathers in flight per row in the lane-per-row path

struct Csr {
    const int* ptr;         // [nrows+1 (+8)]
    const int* idx;         // [nnz (+8)] column indices
    const double* val;      // [nnz (+8)]
    int nrows;
    const int* warp_chunk;  // [W+1] chunk range of each warp of the persistent grid
    const int4* chunk;      // [nchunks] {row0, nnz0, nrows | 0 | -1, nnz}
    int lanes_log2;         // lanes per row in the shared-memory row reduction (1 => reference summation order)
    // rows longer than kChunk ("long rows", nullptr when there are none): cut into pieces, each a chunk of its own
    const int* cta_long;     // [G+1] range of long_rows owned by each CTA
    const int4* long_rows;   // {row, first piece slot, #pieces, 0}
    double* long_part;       // [#pieces] piece sums (scratch)
    int plan_slot;           // 1 + slot of the per-warp plan cache in shared memory (0: not cached)
    // shared-memory-resident copy (batches of small LPs, k_batch: one CTA per problem keeps BOTH matrices of its problem in
    // shared memory for the whole launch: 16-bit column indices); set by the kernel, nullptr otherwise
    const double* sm_val;
    const unsigned short* sm_idx;
    const int* sm_ptr;
};

// Per-warp staging buffer in shared memory: [val window | idx window | row-ptr window].  The value and index windows
// are FIXED-size (kWin elements) and start at the nonzero index (s & ~3) of the chunk, so one chunk (<= kChunk =
// kWin - 4 nonzeros) always fits, every copy has a compile-time size (three TMA bulk copies per chunk), and the inner
// loops use 128-bit shared-memory loads without bounds checks.  Elements of the window outside the chunk belong to
// neighbouring rows (or to the zero padding behind the arrays): their products are computed and never summed.
constexpr int kWin = 256;
static_assert(kChunk + 3 <= kWin && kWin % 128 == 0, "one chunk = one window");
constexpr int kPad = kWin + 8;                                // padding elements behind every matrix array
constexpr int kValWin = kWin * 8;
constexpr int kIdxWin = kWin * 4;
constexpr int kPtrUnits = (kChunkRows + 1 + 3 + 3) / 4;       // 16-byte units: kChunkRows + 1 pointers, <= 3 early
static_assert(kPtrUnits <= 32, "row-pointer window is copied by one instruction");
constexpr int kPtrWin = kPtrUnits * 16;
constexpr int kStageBytes = kValWin + kIdxWin + kPtrWin;
constexpr int kWarpSmemBytes = kStageBytes;
// mbarrier / TMA bulk-copy primitives (the chunk windows are contiguous, 16-byte aligned ranges of the matrix arrays:
// three cp.async.bulk per chunk instead of seven 16-byte cp.async per lane -- the LDGSTS path cost 25 % of all L1
// wavefronts of the kernel, the bulk copies bypass the LSU pipe entirely; SASS: UBLKCP)
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
#ifndef ABIP_TMA_L2_HINT
#define ABIP_TMA_L2_HINT 0  // 1: evict_first, 2: evict_last policy on the matrix stream (experiment, profiles/r02_spmv.md)
#endif
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
#if ABIP_TMA_L2_HINT
    unsigned long long pol;
#if ABIP_TMA_L2_HINT == 1
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
#else
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#endif
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
#else
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_inval(unsigned bar) { asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(bar) : "memory"); }

struct WarpSmem {
    unsigned char* base;  // generic pointer to this warp's buffer
    unsigned base_s;      // same, shared-space address
    unsigned bar_s;       // this warp's mbarrier (shared-space address)
    unsigned parity;      // phase parity of the next wait
    const int4* cur;      // matrix (identified by its chunk array) of the chunk in flight, nullptr: nothing in flight
    int cur_c;            // its chunk index
    int* plan;            // per-warp plan cache: kPlanSlots x {first descriptor (int4), c0, c1, valid, -}

    // this warp's chunk range [c0, c1) of A and the descriptor of its first chunk.  The three dependent global loads
    // cost ~2 us at the start of every phase; the two matrices of the CG loop are cached in shared memory instead.
    __device__ __forceinline__ void get_plan(const Csr& A, int& c0, int& c1, int4& d0) {
        const int gwarp = VB() * kWarps + (threadIdx.x >> 5);
        int* e = plan + 8 * (A.plan_slot - 1);
        if (A.plan_slot > 0) {
            const int4 r = *reinterpret_cast<const int4*>(e + 4);
            if (r.z == 1) {
                c0 = r.x;
                c1 = r.y;
                d0 = *reinterpret_cast<const int4*>(e);
                return;
            }
        }
        c0 = __ldg(A.warp_chunk + gwarp);
        c1 = __ldg(A.warp_chunk + gwarp + 1);
        d0 = make_int4(0, 0, 0, 0);
        if (c0 < c1) d0 = __ldg(A.chunk + c0);
        if (A.plan_slot > 0) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0) {
                *reinterpret_cast<int4*>(e) = d0;
                *reinterpret_cast<int4*>(e + 4) = make_int4(c0, c1, 1, 0);
            }
            __syncwarp();
        }
    }

    // start the asynchronous copy of chunk c (descriptor d) of A into the buffer.  All lanes call it: the buffer was
    // last touched through the generic proxy (LDS/STS of every lane), the bulk copy writes through the async proxy.
    __device__ __forceinline__ void issue(const Csr& A, int c, const int4& d) {
        fence_proxy_async();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            const int si = d.y & ~3, sr = d.x & ~3;
            const bool rows = d.z > 0;
            mbar_expect_tx(bar_s, kValWin + kIdxWin + (rows ? kPtrWin : 0));
            bulk_g2s(base_s, A.val + si, kValWin, bar_s);
            bulk_g2s(base_s + kValWin, A.idx + si, kIdxWin, bar_s);
            if (rows) bulk_g2s(base_s + kValWin + kIdxWin, A.ptr + sr, kPtrWin, bar_s);
        }
        cur = A.chunk;
        cur_c = c;
    }
    __device__ __forceinline__ void wait() {
        mbar_wait(bar_s, parity);
        parity ^= 1u;
    }
    __device__ __forceinline__ void drain() {
        if (cur) wait();
        cur = nullptr;
    }
};

// Deterministic grid-wide reduction helper (see file header).  partials is double-buffered so that a fast
// block starting reduction t+1 can never overwrite values a slow block is still reading for reduction t.
struct Reducer {
    double* partials;  // [2][kMaxRed][G]
    double* sm;        // shared scratch [kMaxRed * kWarps]
    int G;
    int parity;
    WarpSmem ws;       // this warp's SpMV staging slice

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
            partials[(parity * kMaxRed + slot0 + threadIdx.x) * G + VB()] = s;
        }
        __syncthreads();
        // the grid barrier that follows orders these writes before finish()
    }

    // same for max-reductions (inf-norms of the QCP residuals); values must be >= 0
    template <int K>
    __device__ __forceinline__ void block_store_max(double (&v)[K], int slot0 = 0) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double x = v[k];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) x = fmax(x, __shfl_down_sync(0xffffffffu, x, off));
            if (lane == 0) sm[k * kWarps + w] = x;
        }
        __syncthreads();
        if (threadIdx.x < K) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < kWarps; ++i) s = fmax(s, sm[threadIdx.x * kWarps + i]);
            partials[(parity * kMaxRed + slot0 + threadIdx.x) * G + VB()] = s;
        }
        __syncthreads();
    }

    // call after the grid barrier; every block computes the same totals in the same order.
    // MAXMASK: bit k set => slot k is a max-reduction
    template <int K, unsigned MAXMASK = 0u>
    __device__ __forceinline__ void finish(double (&out)[K]) {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int k = w; k < K; k += kWarps) {
            const double* src = partials + (parity * kMaxRed + k) * G;
            const bool mx = (MAXMASK >> k) & 1u;
            double s = 0.0;
            for (int i = lane; i < G; i += 32) {
                const double t = __ldcg(src + i);
                s = mx ? fmax(s, t) : s + t;
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                const double t = __shfl_xor_sync(0xffffffffu, s, off);
                s = mx ? fmax(s, t) : s + t;
            }
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
// K1: CSR SpMV phase ("CSR-stream" per warp, TMA-staged).  fn(row, dot) is called once per row by one thread with
// dot = A[row,:] * x.  For each of its chunks the warp
//   1. waits on its mbarrier for the chunk's bulk copies (issued one chunk ahead, also across grid barriers and
//      across the switch to the matrix of the next phase, `next`);
//   2. L = 1 (short rows): one lane per row -- index, value, gather, multiply, add in the row's own order (the serial
//      summation order of the reference, linsys/common.c:624-634);
//      L > 1: 128-bit loads of indices and values, 8 independent gathers per lane, products back to shared memory,
//      then L lanes per row reduce the rows of the chunk;
//   3. a piece of a long row (d.z <= 0) is reduced by the whole warp into the scratch slot -d.z - 1; after the CTA has
//      finished its chunks one thread per long row adds the pieces in order and calls fn.
// x may have been written earlier in the same kernel (ordinary coherent loads; L1 is invalidated by the grid
// barrier's fence).  Replaces _accum_by_Atrans (reference linsys/common.c:598-639) for both A and A'.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void spmv_prefetch(const Csr& A, WarpSmem& ws) {
    if (A.sm_ptr) return;
    int c0, c1;
    int4 d0;
    ws.get_plan(A, c0, c1, d0);
    if (ws.cur == A.chunk && ws.cur_c == c0) return;
    ws.drain();
    if (c0 < c1) ws.issue(A, c0, d0);
}

#ifdef ABIP_PHASE_TIMING
// debug: SM-clock cycles per warp spent in the sections of the chunk loop, summed over all warps and calls
// [0] wait for the chunk  [1] gather + multiply  [2] row sums + epilogue  [3] issue of the next chunk  [4] chunks
__device__ unsigned long long g_spmv_prof[16];  // [0..7] matrices with 1 lane/row (A'), [8..15] others (A)
#define SPROF_DECL unsigned long long _p0 = clock64(), _pa[4] = {0, 0, 0, 0}, _pn = 0
#define SPROF(i) do { const unsigned long long _t = clock64(); _pa[i] += _t - _p0; _p0 = _t; } while (0)
#define SPROF_RESET() _p0 = clock64()
#define SPROF_FLUSH()                                                                \
    do {                                                                             \
        if ((threadIdx.x & 31) == 0) {                                               \
            const int _b = A.lanes_log2 ? 8 : 0;                                     \
            for (int _i = 0; _i < 4; ++_i) atomicAdd(&g_spmv_prof[_b + _i], _pa[_i]); \
            atomicAdd(&g_spmv_prof[_b + 4], _pn);                                    \
        }                                                                            \
    } while (0)
#else
#define SPROF_DECL do { } while (0)
#define SPROF(i) do { } while (0)
#define SPROF_RESET() do { } while (0)
#define SPROF_FLUSH() do { } while (0)
#endif

// Shared-memory-resident matrix (one CTA owns the whole problem): one thread per row, values / 16-bit indices / row
// pointers out of shared memory, the gathers of a batch of kL1U nonzeros in flight together, the row summed in its own
// order (separate multiply and add, like the lane-per-row path below).
template <class RowFn>
__device__ __forceinline__ void spmv_rows_resident(const Csr& A, const double* x, RowFn fn) {
    for (int row = threadIdx.x; row < A.nrows; row += kBlock) {
        const int a = A.sm_ptr[row], b = A.sm_ptr[row + 1];
        double acc = 0.0;
        for (int k0 = a; k0 < b; k0 += kL1U) {
            double xv[kL1U];
#pragma unroll
            for (int u = 0; u < kL1U; ++u) xv[u] = (k0 + u < b) ? x[A.sm_idx[k0 + u]] : 0.0;
#pragma unroll
            for (int u = 0; u < kL1U; ++u)
                if (k0 + u < b) acc = __dadd_rn(acc, __dmul_rn(A.sm_val[k0 + u], xv[u]));
        }
        fn(row, acc);
    }
}

template <class RowFn>
__device__ __forceinline__ void spmv_rows(const Csr& A, const double* x, WarpSmem& ws, const Csr* next, RowFn fn) {
    if (A.sm_ptr) {  // uniform over the CTA
        spmv_rows_resident(A, x, fn);
        return;
    }
    const int lane = threadIdx.x & 31;
    const int lg = A.lanes_log2;
    const int L = 1 << lg;
    const int sl = lane & (L - 1);
    const int sub = lane >> lg;
    const int rpw = 32 >> lg;
    int c0, c1, nx_c = 0, nx_c1 = 0;
    int4 d, dnx = make_int4(0, 0, 0, 0);  // dnx: first chunk of this warp in the next phase's matrix
    ws.get_plan(A, c0, c1, d);
    if (next) ws.get_plan(*next, nx_c, nx_c1, dnx);
    if (c0 < c1 && !(ws.cur == A.chunk && ws.cur_c == c0)) {  // nothing useful in flight
        ws.drain();
        ws.issue(A, c0, d);
    }
    unsigned char* st = ws.base;
    double* vw = reinterpret_cast<double*>(st);
    const int* iw = reinterpret_cast<const int*>(st + kValWin);
    SPROF_DECL;
    for (int c = c0; c < c1; ++c) {
        int4 dn = dnx;  // descriptor of the chunk to stream in next (loaded now, used after this chunk)
        if (c + 1 < c1) dn = __ldg(A.chunk + c + 1);
        SPROF_RESET();
        ws.wait();
        SPROF(0);
        const int row0 = d.x, s = d.y, nr = d.z, n = d.w;
        const int off = s & 3;
        if (nr > 0 && L == 1) {
            // short rows (mean <= 16 nonzeros): one lane per row straight out of the staged windows -- index, value,
            // gather, multiply, add in the row's own order (the reference's summation order, no FMA contraction so the
            // result does not depend on the path).  No product round trip through shared memory: per 32 nonzeros ~19
            // L1 data-pipe wavefronts instead of ~27, and neighbouring rows gather neighbouring columns together.
            const double* vs = vw + off - s;
            const int* is = iw + off - s;
            const int* rp = reinterpret_cast<const int*>(st + kValWin + kIdxWin) + (row0 & 3);
            // a chunk holds <= 64 rows: lane l takes rows l and l + 32 together, and all gathers of a batch of kL1U
            // nonzeros per row are issued before the first product is formed (a serial index -> gather -> add chain per
            // nonzero was 3/4 of the A' pass: 10 dependent L2 round trips per chunk)
            int a0 = 0, b0 = 0, a1 = 0, b1 = 0;
            if (lane < nr) { a0 = rp[lane]; b0 = rp[lane + 1]; }
            if (lane + 32 < nr) { a1 = rp[lane + 32]; b1 = rp[lane + 33]; }
            const int mx = __reduce_max_sync(0xffffffffu, max(b0 - a0, b1 - a1));
            double acc0 = 0.0, acc1 = 0.0;
            for (int j = 0; j < mx; j += kL1U) {
                double x0[kL1U], x1[kL1U];
#pragma unroll
                for (int u = 0; u < kL1U; ++u) {
                    const int k = a0 + j + u;
                    x0[u] = k < b0 ? x[is[k]] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kL1U; ++u) {
                    const int k = a1 + j + u;
                    x1[u] = k < b1 ? x[is[k]] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kL1U; ++u) {
                    const int k = a0 + j + u;
                    if (k < b0) acc0 = __dadd_rn(acc0, __dmul_rn(vs[k], x0[u]));
                }
#pragma unroll
                for (int u = 0; u < kL1U; ++u) {
                    const int k = a1 + j + u;
                    if (k < b1) acc1 = __dadd_rn(acc1, __dmul_rn(vs[k], x1[u]));
                }
            }
            SPROF(1);
            __syncwarp();
            // the row sums are in registers: stream in the next chunk first, the epilogue's own global loads overlap with it
            if (c + 1 < c1) ws.issue(A, c + 1, dn);
            else if (next && nx_c < nx_c1) ws.issue(*next, nx_c, dn);
            else ws.cur = nullptr;
            d = dn;
            SPROF(3);
            if (lane < nr) fn(row0 + lane, acc0);
            if (lane + 32 < nr) fn(row0 + lane + 32, acc1);
            SPROF(2);
#ifdef ABIP_PHASE_TIMING
            ++_pn;
#endif
            continue;
        }
        double sum0 = 0.0, sum1 = 0.0;
        int epi_rows = 0;
        // 1. gather x for the whole window (8 independent loads per lane), multiply
        double xv[2][4];
        int4 ci[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) ci[u] = reinterpret_cast<const int4*>(iw)[lane + 32 * u];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            xv[u][0] = x[ci[u].x];
            xv[u][1] = x[ci[u].y];
            xv[u][2] = x[ci[u].z];
            xv[u][3] = x[ci[u].w];
        }
        if (nr <= 0) {  // piece of a long row: elements [off, off + n) of the window; slot -nr - 1 of the scratch
            double acc = 0.0;
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = lane + 32 * u;
                const double2 v01 = reinterpret_cast<const double2*>(vw)[2 * q];
                const double2 v23 = reinterpret_cast<const double2*>(vw)[2 * q + 1];
                const int k = 4 * q - off;
                if ((unsigned)k < (unsigned)n) acc = fma(v01.x, xv[u][0], acc);
                if ((unsigned)(k + 1) < (unsigned)n) acc = fma(v01.y, xv[u][1], acc);
                if ((unsigned)(k + 2) < (unsigned)n) acc = fma(v23.x, xv[u][2], acc);
                if ((unsigned)(k + 3) < (unsigned)n) acc = fma(v23.y, xv[u][3], acc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) A.long_part[-nr - 1] = acc;
        } else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int q = lane + 32 * u;
                double2 v01 = reinterpret_cast<const double2*>(vw)[2 * q];
                double2 v23 = reinterpret_cast<const double2*>(vw)[2 * q + 1];
                v01.x *= xv[u][0];
                v01.y *= xv[u][1];
                v23.x *= xv[u][2];
                v23.y *= xv[u][3];
                reinterpret_cast<double2*>(vw)[2 * q] = v01;
                reinterpret_cast<double2*>(vw)[2 * q + 1] = v23;
            }
            __syncwarp();
            SPROF(1);
            // 2. row sums out of shared memory, L lanes per row; each sum is parked in the row's first product slot
            double* vs = vw + off - s;  // vs[k]: product of nonzero k (global nonzero index)
            const int* rp = reinterpret_cast<const int*>(st + kValWin + kIdxWin) + (row0 & 3);
            {  // (L == 1 never gets here: lane-per-row path above)
                for (int base = 0; base < nr; base += rpw) {
                    const int rr = base + sub;
                    const bool ok = rr < nr;
                    double acc = 0.0;
                    int a = 0, b = 0;
                    if (ok) {
                        a = rp[rr];
                        b = rp[rr + 1];
                        for (int k = a + sl; k < b; k += L) acc += vs[k];
                    }
                    for (int o = L >> 1; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o, L);
                    __syncwarp();  // every lane of the group has read its products before slot a is overwritten
                    if (ok && sl == 0 && a < b) vs[a] = acc;
                }
                __syncwarp();
                // 3. the row sums of the chunk (<= 64 rows: lane l holds rows l and l + 32) move to registers ...
                if (lane < nr) { const int a = rp[lane], b = rp[lane + 1]; sum0 = a < b ? vs[a] : 0.0; }
                if (lane + 32 < nr) { const int a = rp[lane + 32], b = rp[lane + 33]; sum1 = a < b ? vs[a] : 0.0; }
                epi_rows = nr;
            }
        }
        __syncwarp();  // every lane is done reading the buffer before it is overwritten
        SPROF(2);
        // ... the next chunk streams in: our own, or (across the grid barrier) the first one of the next phase's matrix ...
        if (c + 1 < c1) ws.issue(A, c + 1, dn);
        else if (next && nx_c < nx_c1) ws.issue(*next, nx_c, dn);
        else ws.cur = nullptr;
        d = dn;
        SPROF(3);
        // ... and the epilogue (one lane per row, its own global loads in flight together) overlaps with that copy
        if (lane < epi_rows) fn(row0 + lane, sum0);
        if (lane + 32 < epi_rows) fn(row0 + lane + 32, sum1);
#ifdef ABIP_PHASE_TIMING
        ++_pn;
#endif
    }
    SPROF_FLUSH();
    if (c0 >= c1 && next && nx_c < nx_c1 && !(ws.cur == next->chunk && ws.cur_c == nx_c)) {
        ws.drain();
        ws.issue(*next, nx_c, dnx);
    }
    // long rows of this CTA: add the piece sums in piece order (pieces were computed by different warps of the CTA)
    if (A.long_rows) {
        const int j0 = __ldg(A.cta_long + VB()), j1 = __ldg(A.cta_long + VB() + 1);
        if (j0 < j1) {  // uniform per CTA
            __syncthreads();
            for (int j = j0 + (int)threadIdx.x; j < j1; j += kBlock) {
                const int4 lr = __ldg(A.long_rows + j);
                double acc = 0.0;
                for (int i = 0; i < lr.z; ++i) acc += A.long_part[lr.y + i];
                fn(lr.x, acc);
            }
            __syncthreads();  // the scratch may be rewritten by the next pass over this matrix
        }
    }
}

// Dynamic shared memory of every persistent kernel: [reducer scratch][mbarriers][plan cache][per-warp staging buffers]
constexpr size_t kRedBytes = sizeof(double) * kMaxRed * kWarps;
constexpr size_t kBarOff = kRedBytes;                                            // one mbarrier per warp
constexpr int kPlanSlots = 2;
constexpr size_t kPlanOff = ((kBarOff + 8 * kWarps + 15) / 16) * 16;             // per-warp plan cache
constexpr size_t kStageOff = ((kPlanOff + 32 * kPlanSlots * kWarps + 127) / 128) * 128;
constexpr size_t kSmemBytes = kStageOff + (size_t)kWarps * kWarpSmemBytes;
__device__ __forceinline__ Reducer make_reducer(unsigned char* smem_raw, double* partials, bool batched = false) {
    vgrid_init(batched);
    const int w = threadIdx.x >> 5;
    Reducer R;
    R.partials = partials;
    R.sm = reinterpret_cast<double*>(smem_raw);
    R.G = VG();
    R.parity = 0;
    unsigned char* wbase = smem_raw + kStageOff + (size_t)w * kWarpSmemBytes;
    R.ws.bar_s = smem_u32(smem_raw + kBarOff + 8 * w);
    R.ws.parity = 0;
    R.ws.plan = reinterpret_cast<int*>(smem_raw + kPlanOff) + 8 * kPlanSlots * w;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < kPlanSlots; ++i) R.ws.plan[8 * i + 6] = 0;
        mbar_init(R.ws.bar_s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
    R.ws.base = wbase;
    R.ws.base_s = smem_u32(wbase);
    R.ws.cur = nullptr;
    R.ws.cur_c = 0;
    return R;
}

// End of one persistent "body": no copy may be outstanding when the CTA exits, and the warp's mbarrier is invalidated so
// that a kernel which runs several bodies in a row (k_batch: device-resident loops) may initialise it again
// (mbarrier.init on a live mbarrier object is undefined).
__device__ __forceinline__ void release_reducer(Reducer& R) {
    R.ws.drain();
    __syncwarp();
    if ((threadIdx.x & 31) == 0) mbar_inval(R.ws.bar_s);
    __syncwarp();
}

// Optional per-phase timing (build with -DABIP_PHASE_TIMING): block 0 / thread 0 accumulates globaltimer deltas per
// phase id into LpCtx::phase_ns (it leaves a barrier only when every block has arrived, so its deltas are the
// critical path of each phase including the barrier).
#ifdef ABIP_PHASE_TIMING
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PHASE_MARK(c, last, id)                                    \
    do {                                                           \
        if (VB() == 0 && threadIdx.x == 0 && (c).phase_ns) { \
            const unsigned long long _t = gtimer();                \
            (c).phase_ns[(id)] += (double)(_t - (last));           \
            (c).phase_ns[16 + (id)] += 1.0;                        \
            (last) = _t;                                           \
        }                                                          \
    } while (0)
#define PHASE_START(last) unsigned long long last = gtimer()
// per-warp busy time of one SpMV phase: phase_ns[32 + which * W + global warp] (which: 0 = A', 1 = A)
#define WARP_T0(t) unsigned long long t = gtimer()
#define WARP_T1(c, t, which)                                                                              \
    do {                                                                                                  \
        if ((threadIdx.x & 31) == 0 && (c).phase_ns)                                                      \
            (c).phase_ns[32 + (which) * VG() * kWarps + VB() * kWarps + (threadIdx.x >> 5)] += \
                (double)(gtimer() - (t));                                                                 \
    } while (0)
#else
#define WARP_T0(t) do { } while (0)
#define WARP_T1(c, t, which) do { } while (0)
#define PHASE_MARK(c, last, id) do { } while (0)
#define PHASE_START(last) do { } while (0)
#endif

// grid barrier; an engine confined to one CTA (batches of small LPs, one CTA per problem) only needs a CTA barrier
__device__ __forceinline__ void grid_sync(cg::grid_group& grid) {
    if (VG() == 1) {
        __threadfence_block();
        __syncthreads();
    } else {
        grid.sync();
    }
}

#define GRID_STRIDE(i, N) \
    for (int i = VB() * kBlock + threadIdx.x, _gs = VG() * kBlock; i < (N); i += _gs)

// ---------------------------------------------------------------------------------------------------------
// Multi-GPU: in-kernel collectives over NVLink peer memory (one process per GPU, buffers shared through CUDA IPC).
// Partition: GPU g owns a block of COLUMNS of A, i.e. a row block of the stored CSR(A'); all
// m-space vectors (y, PCG p/r/Gp, b, D, M) are replicated, n-space vectors (x, s, c, E) are sharded.  The only
// vector exchange is the sum of the partial products A_g x_g (m doubles) -- validate() guarantees m <= n, so this
// is always the smaller vector -- plus a few scalars per ADMM iteration; the PCG scalars need NO communication
// because every GPU computes them redundantly on bit-identical replicated data.
// Every rank sums the G contributions in rank order, so all ranks obtain bit-identical results and take the same
// data-dependent branches.  Flags are sequence numbers: rank r publishes "my contribution #seq is complete" into
// every peer's flag array; buffers are double-buffered by seq parity (a rank can be at most one collective ahead).
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 8;
constexpr int kCommScalars = 16;
struct Comm {
    int G, rank;
    double* vec[kMaxRanks];               // per rank: [2][m_pad] partial m-vectors (IPC-mapped, own buffer included)
    double* red[kMaxRanks];               // per rank: [m_pad] reduced slice (two-step all-reduce, G > 2)
    double* scal[kMaxRanks];              // per rank: [2][kCommScalars]
    unsigned long long* flags[kMaxRanks]; // per rank: [kMaxRanks] "rank q has published collective #"
    unsigned long long* seq;              // local: sequence counter persisted across launches
    int* err;                             // local: sticky error (peer timeout)
    long m_pad;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Device-side state of the communicator inside one kernel (registers).
struct CommState {
    unsigned long long seq;
    bool failed;
};

// publish: local contributions for collective `seq` are complete (caller has executed a grid barrier after the
// last write); wait until every rank has published.  Ends with all threads of the grid released.
__device__ __forceinline__ void comm_exchange(const Comm& cm, CommState& st, cg::grid_group& grid) {
    st.seq += 1;
    if (VB() == 0 && threadIdx.x < cm.G) {
        __threadfence_system();
        st_release_sys(cm.flags[threadIdx.x] + cm.rank, st.seq);
    }
    if (threadIdx.x == 0 && !st.failed) {
        const long long t0 = clock64();
        for (int q = 0; q < cm.G; ++q) {
            while (ld_acquire_sys(cm.flags[cm.rank] + q) < st.seq) {
                if (clock64() - t0 > 20000000000LL) {  // ~10 s: a peer died; fail instead of hanging the GPU
                    *cm.err = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    if (*(volatile int*)cm.err) st.failed = true;
}

// out[i] = sum over ranks of their partial vectors (published in vec[q][parity]), applied through fn(i, sum).
//   G <= 2: every rank reads all copies in rank order (one exchange, (G-1) m doubles over NVLink);
//   G  > 2: reduce-scatter + all-gather (two exchanges, 2 (G-1)/G m doubles): rank r sums its slice of the rows
//           from all peers into its `red` buffer, then every rank collects the G reduced slices.
// Remote copies are read with L1-bypassing loads (peer lines may be stale in the local L1).  Either way each
// element is the same rank-ordered sum on every rank.
template <class Fn>
__device__ __forceinline__ void comm_sum_vec(const Comm& cm, CommState& st, cg::grid_group& grid, int m, Fn fn) {
    grid_sync(grid);  // local partials complete
    comm_exchange(cm, st, grid);
    const long off = (long)(st.seq & 1ull) * cm.m_pad;
    if (cm.G <= 2) {
        GRID_STRIDE(i, m) {
            double s = 0.0;
            for (int q = 0; q < cm.G; ++q) s += __ldcv(cm.vec[q] + off + i);
            fn(i, s);
        }
    } else {
        const int sl = (m + cm.G - 1) / cm.G;
        const int lo = cm.rank * sl, cnt = max(0, min(m, lo + sl) - lo);
        double* mine = cm.red[cm.rank];
        GRID_STRIDE(t, cnt) {
            const int i = lo + t;
            double s = 0.0;
            for (int q = 0; q < cm.G; ++q) s += __ldcv(cm.vec[q] + off + i);
            mine[i] = s;
        }
        // (red is single-buffered: it is written only after the exchange above, i.e. after every rank has finished
        //  all reads of the previous collective)
        grid_sync(grid);
        comm_exchange(cm, st, grid);
        // collect the reduced slices: 4 independent peer loads in flight per thread (peer latency ~ 2-3 us)
        const int gs = VG() * kBlock;
        for (int i0 = VB() * kBlock + threadIdx.x; i0 < m; i0 += 4 * gs) {
            double v4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * gs;
                v4[u] = i < m ? __ldcv(cm.red[i / sl] + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * gs;
                if (i < m) fn(i, v4[u]);
            }
        }
    }
}
__device__ __forceinline__ double* comm_vec_slot(const Comm& cm, const CommState& st) {
    return cm.vec[cm.rank] + (long)((st.seq + 1) & 1ull) * cm.m_pad;  // buffer of the NEXT collective
}

// vals[k] (identical on every thread of this rank) -> sum over ranks, in rank order
template <int K>
__device__ __forceinline__ void comm_sum_scalars(const Comm& cm, CommState& st, cg::grid_group& grid, double (&vals)[K]) {
    static_assert(K <= kCommScalars, "too many scalars");
    double* mine = cm.scal[cm.rank] + ((st.seq + 1) & 1ull) * kCommScalars;
    if (VB() == 0 && threadIdx.x == 0)
        for (int k = 0; k < K; ++k) mine[k] = vals[k];
    grid_sync(grid);
    comm_exchange(cm, st, grid);
    const long off = (long)(st.seq & 1ull) * kCommScalars;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        double s = 0.0;
        for (int q = 0; q < cm.G; ++q) s += __ldcv(cm.scal[q] + off + k);
        vals[k] = s;
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
    double* phase_ns;  // [32] debug phase timing or nullptr
    Comm comm;         // multi-GPU (comm.G == 1: unused)
};

struct SolveOut {
    int its;
    double tol, res;
};


// ---------------------------------------------------------------------------------------------------------
// solve_lin_sys on device (reference linsys/indirect.c:393-434 incl. pcg :321-391 and mat_vec :205-220).
//   b: [m+n] right-hand side, overwritten by the solution;  s: warm start [>= m] or nullptr.
//   EPI: also reduce hdot = sol[0:m+n] . h (epilogue of project_lin_sys, src/abip.c:560); the caller must
//   grid.sync() and R.finish<1>() to obtain it.
// Barriers per solve: 2 + 3 per CG iteration (+1 by the caller).
// ---------------------------------------------------------------------------------------------------------
template <bool EPI, bool DIST>
__device__ __forceinline__ void dev_solve_lin_sys(const LpCtx& c, Reducer& R, cg::grid_group& grid, CommState& cs,
                                                  double* b, const double* s, long iter, SolveOut& out) {
    const int m = c.m;
    double* by = b;
    double* bx = b + m;
    PHASE_START(tl);
    // S1: by += A bx, accumulating |by|^2 of the *incoming* by for the tolerance (indirect.c:406-409, trap 2);
    //     independent of that, tmp = A' s for the warm-start residual.
    double a1[1] = {0.0};
    if constexpr (!DIST) {
        spmv_rows(c.A, bx, R.ws, &c.AT, [&](int row, double a) {
            const double o = by[row];
            a1[0] = fma(o, o, a1[0]);
            by[row] = o + a;
        });
        if (s) spmv_rows(c.AT, s, R.ws, &c.A, [&](int row, double a) { c.tmp[row] = a; });
    } else {  // partial product of the local column block, summed over the GPUs
        double* slot = comm_vec_slot(c.comm, cs);
        spmv_rows(c.A, bx, R.ws, &c.AT, [&](int row, double a) { slot[row] = a; });
        if (s) spmv_rows(c.AT, s, R.ws, &c.A, [&](int row, double a) { c.tmp[row] = a; });
        comm_sum_vec(c.comm, cs, grid, m, [&](int i, double tot) {
            const double o = by[i];
            a1[0] = fma(o, o, a1[0]);
            by[i] = o + tot;
        });
    }
    R.block_store<1>(a1);
    grid_sync(grid);
    R.finish<1>(a1);
    PHASE_MARK(c, tl, 2);
    double tol = sqrt(a1[0]) * (iter < 0 ? 1e-9 : 0.1 / pow((double)iter + 1.0, c.cg_rate));
    tol = fmax(fmax(tol, 1e-7), 1e-9);
    // S2: r = by - (rho s + A tmp), x = s (stored in by), z = M r, p = z   (indirect.c:343-365)
    double a2[2] = {0.0, 0.0};
    if (s) {
        auto warm_epi = [&](int row, double a) {
            const double si = s[row];
            const double ri = by[row] - fma(c.rho_y, si, a);
            const double zi = __ldg(c.M + row) * ri;
            c.r[row] = ri;
            by[row] = si;
            c.p[row] = zi;
            a2[0] = fma(ri, ri, a2[0]);
            a2[1] = fma(zi, ri, a2[1]);
        };
        if constexpr (!DIST) {
            spmv_rows(c.A, c.tmp, R.ws, &c.AT, warm_epi);
        } else {
            double* slot = comm_vec_slot(c.comm, cs);
            spmv_rows(c.A, c.tmp, R.ws, &c.AT, [&](int row, double a) { slot[row] = a; });
            comm_sum_vec(c.comm, cs, grid, m, warm_epi);
        }
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
    grid_sync(grid);
    R.finish<2>(a2);
    PHASE_MARK(c, tl, 3);
    double rn = sqrt(a2[0]);
    double ipzr = a2[1];
    int its = 0;
    if (!(rn < fmin(tol, 1e-18))) {
        // PCG (indirect.c:368-388) with THREE grid barriers per iteration instead of four: the epilogue of the A pass also
        // reduces r.Gp, Gp.Gp, (Mr).Gp, (M Gp).Gp, so that |r - alpha Gp|^2 and (M r').r' -- and with them the stopping
        // test and beta -- are known right after alpha (r' = r - alpha Gp expanded); the updates of x, r and p then run as
        // ONE phase.  The same epilogue re-measures (M r).r and |r|^2 of the current residual, so alpha uses the exact
        // value as in the reference and the expanded forms never accumulate.
        for (int it = 0; it < m; ++it) {
            // L1: tmp = A' p
            WARP_T0(tw1);
            spmv_rows(c.AT, c.p, R.ws, &c.A, [&](int row, double a) { c.tmp[row] = a; });
            WARP_T1(c, tw1, 0);
            grid_sync(grid);
            PHASE_MARK(c, tl, 4);
            // L2: Gp = A tmp + rho p ; seven sums
            double d[7] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            auto gp_epi = [&](int row, double a) {
                const double pi = c.p[row], ri = c.r[row], Mi = __ldg(c.M + row);
                const double gp = fma(c.rho_y, pi, a);
                c.Gp[row] = gp;
                const double zi = Mi * ri, mg = Mi * gp;
                d[0] = fma(pi, gp, d[0]);   // p.Gp
                d[1] = fma(zi, gp, d[1]);   // (M r).Gp
                d[2] = fma(mg, gp, d[2]);   // (M Gp).Gp
                d[3] = fma(ri, gp, d[3]);   // r.Gp
                d[4] = fma(gp, gp, d[4]);   // Gp.Gp
                d[5] = fma(zi, ri, d[5]);   // (M r).r of the current residual
                d[6] = fma(ri, ri, d[6]);   // |r|^2 of the current residual
            };
            if constexpr (!DIST) {
                WARP_T0(tw2);
                spmv_rows(c.A, c.tmp, R.ws, &c.AT, gp_epi);
                WARP_T1(c, tw2, 1);
            } else {
                double* slot = comm_vec_slot(c.comm, cs);
                spmv_rows(c.A, c.tmp, R.ws, &c.AT, [&](int row, double a) { slot[row] = a; });
                comm_sum_vec(c.comm, cs, grid, m, gp_epi);
            }
            R.block_store<7>(d);
            grid_sync(grid);
            R.finish<7>(d);
            PHASE_MARK(c, tl, 5);
            ipzr = d[5];
            const double alpha = ipzr / d[0];
            const double rr_new = fma(alpha, fma(alpha, d[4], -2.0 * d[3]), d[6]);
            const double zr_new = fma(alpha, fma(alpha, d[2], -2.0 * d[1]), ipzr);
            its = it + 1;
            rn = sqrt(fmax(rr_new, 0.0));
            const bool stop = rn < tol;
            const double beta = zr_new / ipzr;
            // L3 + L4: x += alpha p ; r -= alpha Gp ; p = beta p + M r
            GRID_STRIDE(i, m) {
                const double pi = c.p[i];
                by[i] = fma(alpha, pi, by[i]);
                const double ri = fma(-alpha, c.Gp[i], c.r[i]);
                c.r[i] = ri;
                if (!stop) c.p[i] = fma(beta, pi, __ldg(c.M + i) * ri);
            }
            grid_sync(grid);
            PHASE_MARK(c, tl, 6);
            if (stop) break;
        }
    }
    // S4: bx = -bx + A' by   (indirect.c:419-420)
    // (DIST: slot 0 = replicated y part, slot 1 = local x part, summed over the GPUs by the caller)
    double a3[2] = {0.0, 0.0};
    spmv_rows(c.AT, by, R.ws, &c.A, [&](int row, double a) {
        const double nv = a - bx[row];
        bx[row] = nv;
        if (EPI) a3[DIST ? 1 : 0] = fma(nv, __ldg(c.h + m + row), a3[DIST ? 1 : 0]);
    });
    if (EPI) {
        GRID_STRIDE(i, m) a3[0] = fma(by[i], __ldg(c.h + i), a3[0]);
        if constexpr (DIST) R.block_store<2>(a3);
        else {
            double a31[1] = {a3[0]};
            R.block_store<1>(a31);
        }
    }
    PHASE_MARK(c, tl, 8);
    out.its = its;
    out.tol = tol;
    out.res = rn;
}

// finishes the epilogue dot u_t[0:l-1].h started by dev_solve_lin_sys<EPI = true> (call after a grid barrier)
template <bool DIST>
__device__ __forceinline__ double dev_finish_epi_dot(const LpCtx& c, Reducer& R, cg::grid_group& grid, CommState& cs) {
    if constexpr (!DIST) {
        double hd[1];
        R.finish<1>(hd);
        return hd[0];
    } else {
        double hd[2];
        R.finish<2>(hd);
        double x[1] = {hd[1]};
        comm_sum_scalars<1>(c.comm, cs, grid, x);
        return hd[0] + x[0];
    }
}

// ---------------------------------------------------------------------------------------------------------
// Right-hand side of project_lin_sys (reference src/abip.c:551-558, src/adaptive.c:91-96):
//   ut = u + v; ut[0:m] *= rho_y; ut[0:l-1] -= ut[l-1] h; ut[0:l-1] -= h (ut.g)/(g_th+1); ut[m:l-1] *= -1
// Optionally records u_prev = u on the (x,tau) tail (abip.c:2133; only the tail is ever read back).
// Ends with a grid barrier.
// ---------------------------------------------------------------------------------------------------------
template <bool DIST>
__device__ __forceinline__ void dev_build_rhs(const LpCtx& c, Reducer& R, cg::grid_group& grid, CommState& cs,
                                              const double* u, const double* v, double* ut, double* u_prev_out) {
    const int m = c.m, lm1 = c.m + c.n;
    spmv_prefetch(c.A, R.ws);  // the solve that follows starts with A; its first chunks stream in meanwhile
    const double tt = u[lm1] + v[lm1];
    double a[2] = {0.0, 0.0};  // DIST: [replicated y part, local x part]
    GRID_STRIDE(i, lm1) {
        const double ui = u[i];
        double w = ui + v[i];
        if (i < m) w *= c.rho_y;
        else if (u_prev_out) u_prev_out[i] = ui;
        w = fma(-tt, __ldg(c.h + i), w);
        ut[i] = w;
        const int slot = (DIST && i >= m) ? 1 : 0;
        a[slot] = fma(w, c.g[i], a[slot]);  // (plain load: a batch problem computes g earlier in the same launch)
    }
    if (VB() == 0 && threadIdx.x == 0) {
        ut[lm1] = tt;
        if (u_prev_out) u_prev_out[lm1] = u[lm1];
    }
    double dot;
    if constexpr (!DIST) {
        double a1[1] = {a[0]};
        R.block_store<1>(a1);
        grid_sync(grid);
        R.finish<1>(a1);
        dot = a1[0];
    } else {
        R.block_store<2>(a);
        grid_sync(grid);
        R.finish<2>(a);
        double x[1] = {a[1]};
        comm_sum_scalars<1>(c.comm, cs, grid, x);
        dot = a[0] + x[0];
    }
    const double coef = -dot / (c.g_th + 1.0);
    GRID_STRIDE(i, lm1) {
        double w = fma(coef, __ldg(c.h + i), ut[i]);
        if (i >= m) w = -w;
        ut[i] = w;
    }
    grid_sync(grid);
}

// barrier proximal step on one coordinate (src/abip.c:742-746)
__device__ __forceinline__ double barrier_prox(double t, double lam) {
    const double hlf = t / 2;
    return hlf + sqrt(fma(hlf, hlf, lam));
}
