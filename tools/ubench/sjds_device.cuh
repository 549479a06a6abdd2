// sjds_device.cuh -- device side of the SpMV of the engine: sliced-JDS rows fed by per-warp TMA rings (sm_100a).
//
// Layout and plan: sjds_host.h.  One warp owns a list of chunks (runs of whole steps of one 32-row slice, or one piece of
// a long row; <= kCH elements).  The chunk's contiguous ranges of the value / index arrays (+ the slice's 64-byte lane
// table) are streamed into the warp's ring of kNS shared-memory stages by TMA bulk copies (cp.async.bulk + one mbarrier
// per stage, SASS UBLKCP), kNS chunks ahead -- also across grid barriers into the first chunks of the next phase's
// matrix -- so the only latency a warp ever waits for is that of its gathers.  Inside a stage the nonzeros are
// step-major: lane l reads element (offset of the step) + l, i.e. conflict-free 64-bit / 32-bit shared-memory loads, no
// row pointers, no product round trip, and every row is summed by one lane in the row's own order (separate multiply
// and add: the serial summation order of the reference, linsys/common.c:624-634).
// Replaces _accum_by_Atrans (reference linsys/common.c:598-639) for both A and A'.
#pragma once
#include <cuda_runtime.h>

#ifndef ABIP_CH
#define ABIP_CH 256   // elements per stage (must equal sjds::kChunkElems of the plan)
#endif
#ifndef ABIP_NS
#define ABIP_NS 2     // stages per warp
#endif
#ifndef ABIP_GU
#define ABIP_GU 8     // gathers in flight per lane
#endif
constexpr int kCH = ABIP_CH;
constexpr int kNS = ABIP_NS;
constexpr int kValBytes = kCH * 8;
constexpr int kIdxBytes = kCH * 4;
constexpr int kMetaOff = kValBytes + kIdxBytes;       // 32 x u16 lane table of the slice
constexpr int kDescOff = kMetaOff + 64;               // int4 descriptor of the chunk held by the stage
constexpr int kStageBytes = ((kDescOff + 16 + 127) / 128) * 128;
constexpr int kInfoLastBit = 1 << 16, kInfoLongBit = 1 << 18;
constexpr int kSkipLaneDev = 0xff;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
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
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One matrix in SJDS form + its plan (device view)
struct Sjds {
    const double* val;             // [stored + pad]
    const int* idx;                // [stored + pad]
    const unsigned short* meta;    // [32 * nslices] length | source lane << 8
    const int4* chunk;             // {first element, #elements, first row of the slice | piece slot, info}, grouped by warp
    const int* warp_chunk;         // [W + 1]
    const int4* long_rows;         // {row, first piece slot, #pieces, 0}, grouped by CTA (nullptr: no long rows)
    const int* cta_long;           // [G + 1]
    double* long_part;             // [#pieces] piece sums (scratch)
    int nrows;
    int plan_slot;                 // 1 + slot of the per-warp plan cache in shared memory (0: not cached)
};

// Per-warp TMA ring
struct WarpRing {
    unsigned char* base;  // this warp's stages (generic pointer)
    unsigned base_s;      // same, shared-space address
    unsigned bar_s;       // mbarrier of stage i at bar_s + 8 i
    unsigned par;         // bit i: parity of the next wait on stage i
    int head;             // stage holding the oldest chunk in flight
    int nfl;              // chunks in flight of the matrix `cur`: indices cur_c .. cur_c + nfl - 1
    const int4* cur;      // identifies the matrix (by its chunk array); nullptr: nothing in flight
    int cur_c;
    int* plan;            // per-warp plan cache in shared memory: kPlanSlots x {c0, c1, valid, -}

    __device__ __forceinline__ void get_plan(const Sjds& A, int gwarp, int& c0, int& c1) {
        int* e = plan + 4 * (A.plan_slot - 1);
        if (A.plan_slot > 0) {
            const int4 r = *reinterpret_cast<const int4*>(e);
            if (r.z == 1) { c0 = r.x; c1 = r.y; return; }
        }
        c0 = __ldg(A.warp_chunk + gwarp);
        c1 = __ldg(A.warp_chunk + gwarp + 1);
        if (A.plan_slot > 0) {
            __syncwarp();
            if ((threadIdx.x & 31) == 0) *reinterpret_cast<int4*>(e) = make_int4(c0, c1, 1, 0);
            __syncwarp();
        }
    }
    // start the copies of chunk d of A into stage `slot` (= head + number of chunks in flight).  All lanes call it: the
    // stage was last read through the generic proxy, the bulk copies write through the async proxy.
    __device__ __forceinline__ void issue(const Sjds& A, const int4& d, int slot) {
        fence_proxy_async();
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            const int st = slot >= kNS ? slot - kNS : slot;
            unsigned char* sp = base + st * kStageBytes;
            const unsigned ss = base_s + st * kStageBytes, bar = bar_s + 8 * st;
            *reinterpret_cast<int4*>(sp + kDescOff) = d;
            const int s4 = d.x & ~3;
            const int ne = ((d.x & 3) + d.y + 3) & ~3;
            const bool slice = !(d.w & kInfoLongBit);
            mbar_expect_tx(bar, ne * 12 + (slice ? 64 : 0));
            if (ne) {
                bulk_g2s(ss, A.val + s4, ne * 8, bar);
                bulk_g2s(ss + kValBytes, A.idx + s4, ne * 4, bar);
            }
            if (slice) bulk_g2s(ss + kMetaOff, A.meta + d.z, 64, bar);
        }
        __syncwarp();
    }
    __device__ __forceinline__ void wait_head() {
        mbar_wait(bar_s + 8 * head, (par >> head) & 1u);
        par ^= 1u << head;
    }
    __device__ __forceinline__ void pop() {
        head = head + 1 == kNS ? 0 : head + 1;
        --nfl;
        ++cur_c;
    }
    __device__ __forceinline__ void drain() {
        while (nfl > 0) { wait_head(); pop(); }
        cur = nullptr;
    }
    // make the ring hold chunks c0.. of A (as many as fit)
    __device__ __forceinline__ void prime(const Sjds& A, int c0, int c1) {
        if (!(cur == A.chunk && cur_c == c0)) {
            drain();
            cur = A.chunk;
            cur_c = c0;
        }
        while (nfl < kNS && cur_c + nfl < c1) {
            const int4 d = __ldg(A.chunk + cur_c + nfl);
            issue(A, d, head + nfl);
            ++nfl;
        }
    }
};

__device__ __forceinline__ void sjds_prefetch(const Sjds& A, WarpRing& rg, int gwarp) {
    int c0, c1;
    rg.get_plan(A, gwarp, c0, c1);
    rg.prime(A, c0, c1);
}

// fn(row, dot) is called once per row by one thread with dot = A[row,:] * x.  `next`: matrix of the phase that follows
// (its first chunks are streamed in while this phase drains).  x may have been written earlier in the same kernel.
// Contains CTA barriers when the matrix has long rows (uniform per CTA).
template <int KBLOCK, class RowFn>
__device__ __forceinline__ void sjds_rows(const Sjds& A, const double* x, WarpRing& rg, const Sjds* next, int vblock,
                                          RowFn fn) {
    constexpr int U = ABIP_GU;
    const int lane = threadIdx.x & 31;
    const int gwarp = vblock * (KBLOCK / 32) + (threadIdx.x >> 5);
    int c0, c1, n0 = 0, n1 = 0;
    rg.get_plan(A, gwarp, c0, c1);
    if (next) rg.get_plan(*next, gwarp, n0, n1);
    rg.prime(A, c0, c1);
    int nfl_next = 0;
    double acc = 0.0;
    for (int c = c0; c < c1; ++c) {
        const int cn = c + kNS;
        int4 dn = make_int4(0, 0, 0, 0);
        if (cn < c1) dn = __ldg(A.chunk + cn);
        else if (next && n0 + nfl_next < n1) dn = __ldg(next->chunk + n0 + nfl_next);
        rg.wait_head();
        const unsigned char* st = rg.base + rg.head * kStageBytes;
        const int4 d = *reinterpret_cast<const int4*>(st + kDescOff);
        const double* sv = reinterpret_cast<const double*>(st) + (d.x & 3);
        const int* si = reinterpret_cast<const int*>(st + kValBytes) + (d.x & 3);
        if (d.w & kInfoLongBit) {
            double a = 0.0;
#pragma unroll
            for (int k0 = 0; k0 < kCH; k0 += 32 * U) {
                double xx[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = k0 + 32 * u + lane;
                    xx[u] = k < d.y ? x[si[k]] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int k = k0 + 32 * u + lane;
                    a = fma(k < d.y ? sv[k] : 0.0, xx[u], a);
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0) A.long_part[d.z] = a;
        } else {
            const unsigned mt = reinterpret_cast<const unsigned short*>(st + kMetaOff)[lane];
            const int len = mt & 0xff, src = mt >> 8;
            const int j0 = (d.w >> 8) & 0xff, jend = j0 + (d.w & 0xff);
            const int lim = min(len, jend);
            int off = lane;
            for (int j = j0; j < jend; j += U) {
                double xx[U];
                int oo[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const bool act = j + u < lim;
                    const unsigned mk = __ballot_sync(0xffffffffu, act);
                    oo[u] = off;
                    xx[u] = act ? x[si[off]] : 0.0;
                    off += __popc(mk);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const double v = (j + u < lim) ? sv[oo[u]] : 0.0;
                    acc = __dadd_rn(acc, __dmul_rn(v, xx[u]));
                }
            }
            if (d.w & kInfoLastBit) {
                if (src != kSkipLaneDev) fn(d.z + src, acc);
                acc = 0.0;
            }
        }
        __syncwarp();  // every lane is done reading the stage before it is overwritten
        rg.pop();
        if (cn < c1) {
            rg.issue(A, dn, rg.head + rg.nfl + nfl_next);
            ++rg.nfl;
        } else if (next && n0 + nfl_next < n1) {
            rg.issue(*next, dn, rg.head + rg.nfl + nfl_next);
            ++nfl_next;
        }
    }
    if (next && n0 < n1) {
        // hand the ring over to the next phase's matrix (top up when this warp had fewer than kNS chunks of its own)
        rg.cur = next->chunk;
        rg.cur_c = n0;
        rg.nfl = nfl_next;
        rg.prime(*next, n0, n1);
    } else {
        rg.cur = nullptr;
    }
    // long rows of this CTA: add the piece sums in piece order (pieces were computed by different warps of the CTA)
    if (A.long_rows) {
        const int j0 = __ldg(A.cta_long + vblock), j1 = __ldg(A.cta_long + vblock + 1);
        if (j0 < j1) {  // uniform per CTA
            __syncthreads();
            for (int j = j0 + (int)threadIdx.x; j < j1; j += KBLOCK) {
                const int4 lr = __ldg(A.long_rows + j);
                double a = 0.0;
                for (int i = 0; i < lr.z; ++i) a += A.long_part[lr.y + i];
                fn(lr.x, a);
            }
            __syncthreads();  // the scratch may be rewritten by the next pass over this matrix
        }
    }
}

// shared-memory carve-out of the rings: [mbarriers kWarps x kNS][plan cache][stages]
constexpr int kPlanSlots = 2;
template <int KWARPS>
struct RingLayout {
    static constexpr size_t bar_bytes = 8 * KWARPS * kNS;
    static constexpr size_t plan_off = ((bar_bytes + 15) / 16) * 16;
    static constexpr size_t stage_off = ((plan_off + 16 * kPlanSlots * KWARPS + 127) / 128) * 128;
    static constexpr size_t bytes = stage_off + (size_t)KWARPS * kNS * kStageBytes;
};
template <int KWARPS>
__device__ __forceinline__ WarpRing make_ring(unsigned char* smem /* 128-byte aligned */) {
    using L = RingLayout<KWARPS>;
    const int w = threadIdx.x >> 5;
    WarpRing rg;
    rg.bar_s = smem_u32(smem + 8 * kNS * w);
    rg.plan = reinterpret_cast<int*>(smem + L::plan_off) + 4 * kPlanSlots * w;
    rg.base = smem + L::stage_off + (size_t)w * kNS * kStageBytes;
    rg.base_s = smem_u32(rg.base);
    rg.par = 0;
    rg.head = 0;
    rg.nfl = 0;
    rg.cur = nullptr;
    rg.cur_c = 0;
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < kPlanSlots; ++i) rg.plan[4 * i + 2] = 0;
#pragma unroll
        for (int i = 0; i < kNS; ++i) mbar_init(rg.bar_s + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_proxy_async();
    __syncwarp();
    return rg;
}
