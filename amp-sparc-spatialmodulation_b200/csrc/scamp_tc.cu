// SCAMP's two batched complex GEMMs (scamp.py:48,56) on the 5th-generation tensor cores: tcgen05.mma kind::tf32 with the
// 3xTF32 split (hi*hi + hi*lo + lo*hi, float32 accumulation in TMEM), operands staged in shared memory in the canonical
// K-major no-swizzle UMMA layout, accumulators read back with tcgen05.ld for the fused epilogues of scamp.cu.
//
//   MODE 0 (residual) : S[f][i] = sum_j Xh[f][j] A[i][j]         B-source = A   [n][N]
//   MODE 1 (estimate) : S[f][j] = sum_i Zs[f][i] conj(A[i][j])   B-source = A^T [N][n]  (transposed once per call)
//
// One CTA = one 128-frame x 64-output tile.  Complex arithmetic on split planes: the frame tile (MMA operand A, M = 128)
// and the design-matrix tile (MMA operand B, N = 64) are each split into {re,im} x {hi,lo} float32 planes while they are
// staged (hi = value with the low 13 mantissa bits cleared -- what kind::tf32 reads anyway --, lo = value - hi, exact);
// per K = 8 step twelve MMAs feed two accumulators
//   D_re += Xr Br -/+ Xi Bi        D_im += Xr Bi +/- ... Xi Br      (signs by MODE; the minus is the descriptor's negate-A bit)
// 32 x 32 tiles of the design matrix that are entirely zero (band structure, channel.py:89-91) are skipped through the
// tile map of scamp.cu; tiles whose 128 frames all met the exit test are skipped altogether.
// Shared memory is single-buffered (96 KiB per CTA): two CTAs per SM overlap one's staging with the other's MMAs.
#include <cstdlib>

#include "blockops.cuh"
#include "kernels.h"
#include "scamp_ws.cuh"

namespace ampsm {

namespace {

constexpr int TM = 128;          // frames per tile (UMMA M)
constexpr int kMaxKBlocks = 1024; // K blocks per tile the non-zero bitmap can hold (K <= 32768)
constexpr int TK = 32;           // reduction elements per staged block (4 MMA K-steps of 8)
constexpr int kTcThreads = 256;
constexpr int kTmemCols = 128;   // D_re | D_im, up to 64 float32 columns each

constexpr int kXPlane = TM * TK * 4;         // bytes of one frame-tile plane (16 KiB)
constexpr int kSmemX = 0;                    // planes: Xr_hi, Xr_lo, Xi_hi, Xi_lo
constexpr int kSmemB = 4 * kXPlane;          // planes: Br_hi, Br_lo, Bi_hi, Bi_lo (TN * TK * 4 bytes each)
template <int TN>
struct TcSmem {
    static constexpr int bplane = TN * TK * 4;
    static constexpr int bar = kSmemB + 4 * bplane;
    static constexpr int total = bar + 64;
};

// K-major, no swizzle: 8 rows x 16 bytes core matrices; a plane is [K chunk of 4][row group of 8][8 rows][16 B]
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);             // start address            bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;       // leading byte offset      bits [16,30)  (next K chunk)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;       // stride byte offset       bits [32,46)  (next 8-row group)
    d |= (uint64_t)1 << 46;                                  // descriptor version 1 (Blackwell)
    return d;                                                // base offset 0, layout type SWIZZLE_NONE
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = TN, optional negate-A
template <int TN>
__device__ __forceinline__ constexpr uint32_t umma_idesc(bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
        "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}

// split four values into the hi / lo planes: 16 bytes each at (chunk, row) of the canonical layout
__device__ __forceinline__ void split_store(unsigned char* hi_plane, unsigned char* lo_plane, uint32_t off, float a, float b, float c,
                                            float d) {
    const float ah = __uint_as_float(__float_as_uint(a) & 0xffffe000u), bh = __uint_as_float(__float_as_uint(b) & 0xffffe000u);
    const float ch = __uint_as_float(__float_as_uint(c) & 0xffffe000u), dh = __uint_as_float(__float_as_uint(d) & 0xffffe000u);
    *reinterpret_cast<float4*>(hi_plane + off) = make_float4(ah, bh, ch, dh);
    *reinterpret_cast<float4*>(lo_plane + off) = make_float4(a - ah, b - bh, c - ch, d - dh);
}

// fused update of one output tile from the complex sums in shared memory: consecutive threads on consecutive outputs (every
// global access of a warp is one contiguous run); eight elements per thread at a time, loads issued before the first use
template <int MODE, int TN>
__device__ __forceinline__ void tc_epilogue(const ScampWs& w, const Geom& g, const float2* __restrict__ y, long long F, const float2* tile,
                                            const int* act_s, long long f0, int o0, int Odim, int tid) {
    // eight elements per thread at a time: all global loads of a batch are issued before the first use
    constexpr int kEpiBatch = 8;
    for (int e0 = tid; e0 < TM * TN; e0 += kTcThreads * kEpiBatch) {
        float2 in0[kEpiBatch], in1[kEpiBatch];
        float sc0[kEpiBatch], sc1[kEpiBatch];
        bool ok[kEpiBatch];
#pragma unroll
        for (int u = 0; u < kEpiBatch; ++u) {
            const int e = e0 + u * kTcThreads;
            const int fr = e / TN, q = e % TN;
            const long long f = f0 + fr;
            const int o = o0 + q;
            ok[u] = e < TM * TN && f < F && o < Odim && act_s[fr];
            in0[u] = in1[u] = make_float2(0.f, 0.f);
            sc0[u] = sc1[u] = 0.f;
            if (ok[u]) {
                if (MODE == 0) {
                    const int blk = o / g.Nr;                          // row block (Mr = Nr)
                    in0[u] = y[f * g.n + o];
                    in1[u] = w.Z[f * g.n + o];
                    sc0[u] = w.b[f * g.Lout + blk];
                    sc1[u] = w.phi[f * g.Lout + blk];
                } else {
                    in0[u] = w.Xh[f * g.N + o];
                    sc0[u] = w.tau[f * g.Lin + o / g.Nt];              // column block (Mc = Nt)
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kEpiBatch; ++u) {
            if (!ok[u]) continue;
            const int e = e0 + u * kTcThreads;
            const int fr = e / TN, q = e % TN;
            const long long f = f0 + fr;
            const int o = o0 + q;
            const float2 s = tile[fr * (TN + 1) + q];
            if (MODE == 0) {
                const float2 zn = make_float2(in0[u].x - s.x + sc0[u] * in1[u].x, in0[u].y - s.y + sc0[u] * in1[u].y);   // scamp.py:48
                w.Z[f * g.n + o] = zn;
                w.Zs[f * g.n + o] = cdiv_real(zn, sc1[u]);                                                              // z / phi_use
            } else {
                w.Xmap[f * g.N + o] = make_float2(fmaf(sc0[u], s.x, in0[u].x), fmaf(sc0[u], s.y, in0[u].y));             // scamp.py:56
            }
        }
    }
}

// Bm: the B-source matrix [Odim][Kdim] complex64 row-major (A for MODE 0, A^T for MODE 1)
template <int MODE, int TN>
__global__ void __launch_bounds__(kTcThreads, 2) scamp_tc_gemm_kernel(ScampWs w, Geom g, const float2* __restrict__ Bm,
                                                                      const float2* __restrict__ y, long long F) {
    constexpr int kBPlane = TcSmem<TN>::bplane;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ int act_s[TM];
    __shared__ uint32_t nzbits[kMaxKBlocks / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + TcSmem<TN>::bar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long f0 = (long long)blockIdx.y * TM;
    const int o0 = blockIdx.x * TN;                       // first output of this tile
    const int Kdim = MODE == 0 ? g.N : g.n;               // reduction length
    const int Odim = MODE == 0 ? g.n : g.N;               // output width
    const float2* __restrict__ Xin = MODE == 0 ? w.Xh : w.Zs;

    int any_active = 0;
    if (tid < TM) {
        const long long f = f0 + tid;
        const int a = (f < F) ? w.active[f] : 0;
        act_s[tid] = a;
        any_active = a;
    }
    if (!__syncthreads_or(any_active)) return;            // before any TMEM allocation

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    unsigned char* Xp[4] = {smem + kSmemX, smem + kSmemX + kXPlane, smem + kSmemX + 2 * kXPlane, smem + kSmemX + 3 * kXPlane};
    unsigned char* Bp[4] = {smem + kSmemB, smem + kSmemB + kBPlane, smem + kSmemB + 2 * kBPlane, smem + kSmemB + 3 * kBPlane};
    constexpr uint32_t kXLbo = TM * 16, kBLbo = TN * 16, kSbo = 128;

    constexpr int kXItems = TM * (TK / 4) / kTcThreads;    // (row, K chunk) items of the frame tile per thread   (4)
    constexpr int kBItems = TN * (TK / 4) / kTcThreads;    // ... of the design tile                              (2)
    static_assert(TM * (TK / 4) % kTcThreads == 0 && TN * (TK / 4) % kTcThreads == 0, "items must divide over the threads");
    static_assert(TN == 32 || TN == 64, "epilogue mapping");
    const bool vec_ok = (Kdim & 1) == 0;                   // 16-byte aligned rows
    // which K blocks hold a non-zero 32 x 32 tile of A for this output tile: one K block per thread, a bitmap in shared memory
    // (a serial scan would chain hundreds of dependent global loads in front of the first MMA)
    const int nkb = (Kdim + TK - 1) / TK;
    for (int kb0 = 0; kb0 < nkb; kb0 += kTcThreads) {
        const int kb = kb0 + tid;
        bool nz = false;
        if (kb < nkb) {
#pragma unroll
            for (int hh = 0; hh < (TN + TILE - 1) / TILE; ++hh) {
                const int ot = (o0 + hh * TILE) / TILE;
                if (o0 + hh * TILE < Odim) nz |= (MODE == 0 ? w.nz[ot * w.nzc + kb] : w.nz[kb * w.nzc + ot]) != 0;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) nzbits[(kb0 >> 5) + warp] = m;
    }
    __syncthreads();
    auto next_block = [&](int k0) {                        // first non-zero K block at or after k0 (Kdim if none)
        int kb = k0 / TK;
        while (kb < nkb) {
            const uint32_t wbits = nzbits[kb >> 5] >> (kb & 31);
            if (wbits) return (kb + __ffs(wbits) - 1) * TK;
            kb = (kb | 31) + 1;
        }
        return Kdim;
    };
    // global memory -> registers (issued BEFORE waiting for the previous block's MMAs: the loads fly under them)
    float4 xv[kXItems][2];
    float2 bv[kBItems][4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int u = 0; u < kXItems; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TM, c = it / TM;
            const long long f = f0 + r;
            const int k = k0 + c * 4;
            xv[u][0] = xv[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (f < F && k + 3 < Kdim && vec_ok) {
                const float4* src = reinterpret_cast<const float4*>(Xin + f * Kdim + k);
                xv[u][0] = src[0];
                xv[u][1] = src[1];
            } else if (f < F) {
                float2 e[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) e[q] = (k + q < Kdim) ? Xin[f * Kdim + k + q] : make_float2(0.f, 0.f);
                xv[u][0] = make_float4(e[0].x, e[0].y, e[1].x, e[1].y);
                xv[u][1] = make_float4(e[2].x, e[2].y, e[3].x, e[3].y);
            }
        }
#pragma unroll
        for (int u = 0; u < kBItems; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TN, c = it / TN;
            const int o = o0 + r;
            const int k = k0 + c * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) bv[u][q] = (o < Odim && k + q < Kdim) ? __ldg(Bm + (size_t)o * Kdim + k + q) : make_float2(0.f, 0.f);
        }
    };
    // registers -> the eight hi / lo planes in the canonical layout
    auto stage = [&]() {
#pragma unroll
        for (int u = 0; u < kXItems; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TM, c = it / TM;
            const uint32_t off = (uint32_t)c * kXLbo + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16;
            split_store(Xp[0], Xp[1], off, xv[u][0].x, xv[u][0].z, xv[u][1].x, xv[u][1].z);      // real parts of the 4 elements
            split_store(Xp[2], Xp[3], off, xv[u][0].y, xv[u][0].w, xv[u][1].y, xv[u][1].w);      // imaginary parts
        }
#pragma unroll
        for (int u = 0; u < kBItems; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TN, c = it / TN;
            const uint32_t off = (uint32_t)c * kBLbo + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16;
            split_store(Bp[0], Bp[1], off, bv[u][0].x, bv[u][1].x, bv[u][2].x, bv[u][3].x);
            split_store(Bp[2], Bp[3], off, bv[u][0].y, bv[u][1].y, bv[u][2].y, bv[u][3].y);
        }
    };

    // shared-memory descriptors of the eight planes (K step j adds 2 j LBO to the start-address field)
    const uint32_t d_re = tmem_base, d_im = tmem_base + TN;
    uint64_t xd[4], bd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        xd[q] = umma_desc(smem_u32(Xp[q]), kXLbo, kSbo);
        bd[q] = umma_desc(smem_u32(Bp[q]), kBLbo, kSbo);
    }
    uint32_t phase = 0;
    int blocks_done = 0;
    int k0 = next_block(0);
    if (k0 < Kdim) fetch(k0);
    while (k0 < Kdim) {
        if (blocks_done > 0) {                             // the previous block's MMAs still read the planes
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        stage();
        fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core (async proxy)
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < TK / 8; ++j) {
                const uint64_t xo = (uint64_t)((2 * j) * kXLbo >> 4), bo = (uint64_t)((2 * j) * kBLbo >> 4);   // start-address field
                const uint32_t acc = (blocks_done > 0 || j > 0) ? 1u : 0u;
                // (x plane pair, b plane pair, accumulator, negate) for the four real products of the complex product
                //   MODE 0: re = XrBr - XiBi, im = XrBi + XiBr;   MODE 1 (conj): re = XrBr + XiBi, im = XiBr - XrBi
                const int xs[4] = {0, 2, 0, 2}, bs[4] = {0, 2, 2, 0};
                const bool ng[4] = {false, MODE == 0, MODE == 1, false};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t dst = (q < 2) ? d_re : d_im;
                    const uint32_t accq = (q == 0 || q == 2) ? acc : 1u;
                    const uint32_t id = ng[q] ? umma_idesc<TN>(true) : umma_idesc<TN>(false);
                    // 3xTF32: hi*hi + hi*lo + lo*hi
                    umma_tf32(dst, xd[xs[q]] + xo, bd[bs[q]] + bo, id, accq);
                    umma_tf32(dst, xd[xs[q]] + xo, bd[bs[q] + 1] + bo, id, 1u);
                    umma_tf32(dst, xd[xs[q] + 1] + xo, bd[bs[q]] + bo, id, 1u);
                }
            }
            umma_commit(bar);                              // arrives when every MMA issued so far has completed
        }
        ++blocks_done;
        k0 = next_block(k0 + TK);
        if (k0 < Kdim) fetch(k0);                          // next block's loads fly under this block's MMAs
    }
    if (blocks_done > 0) {
        mbar_wait(bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    // ---- epilogue.  TMEM -> registers (thread = one frame = one TMEM lane; warps 0-3 take outputs 0-31, warps 4-7 outputs
    // 32-63) -> shared memory (the operand planes are free now), then the fused update with consecutive threads on
    // consecutive outputs: every global access of a warp is one contiguous 256-byte run
    float2* tile = reinterpret_cast<float2*>(smem);            // [TM][TN + 1] complex sums
    if (warp < 4 * (TN / 32)) {
        const int fr = (warp & 3) * 32 + lane;
        const int half = warp >> 2;
        uint32_t vr[32], vi[32];
        if (blocks_done > 0) {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
            tmem_ld32(taddr, vr);
            tmem_ld32(taddr + TN, vi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) vr[q] = vi[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) tile[fr * (TN + 1) + half * 32 + q] = make_float2(__uint_as_float(vr[q]), __uint_as_float(vi[q]));
    }
    __syncthreads();
    tc_epilogue<MODE, TN>(w, g, y, F, tile, act_s, f0, o0, Odim, tid);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}


// ---- variant for long reductions (the residual GEMM: K = N columns, ~64 non-zero blocks per tile): the raw complex tiles
// arrive through a two-stage cp.async ring (one CTA per SM; the next block is requested before the current one is
// converted, so its loads fly under the conversion and the MMAs), so the global-load latency that
// bounds the register-prefetch kernel above (ncu: 53 % of its samples wait on the first use of the prefetched
// registers) is covered; the split into hi / lo planes reads the ring from shared memory.
constexpr int kRingStages = 2;
constexpr int kRawRow = TK * 8 + 16;                      // bytes per raw row: 32 complex + 16 pad (conflict-free 16-byte columns)
template <int TN>
struct RingSmem {
    static constexpr int bplane = TN * TK * 4;
    static constexpr int planes = 4 * kXPlane + 4 * bplane;
    static constexpr int stage = (TM + TN) * kRawRow;
    static constexpr int ring = planes;
    static constexpr int bar = ring + kRingStages * stage;
    static constexpr int total = bar + 64;
};
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}

template <int MODE, int TN>
__global__ void __launch_bounds__(kTcThreads, 1) scamp_tc_ring_kernel(ScampWs w, Geom g, const float2* __restrict__ Bm,
                                                                      const float2* __restrict__ y, long long F) {
    using RS = RingSmem<TN>;
    constexpr int kBPlane = RS::bplane;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ int act_s[TM];
    __shared__ uint32_t nzbits[kMaxKBlocks / 32];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + RS::bar);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long f0 = (long long)blockIdx.y * TM;
    const int o0 = blockIdx.x * TN;
    const int Kdim = MODE == 0 ? g.N : g.n;
    const int Odim = MODE == 0 ? g.n : g.N;
    const float2* __restrict__ Xin = MODE == 0 ? w.Xh : w.Zs;

    int any_active = 0;
    if (tid < TM) {
        const long long f = f0 + tid;
        const int a = (f < F) ? w.active[f] : 0;
        act_s[tid] = a;
        any_active = a;
    }
    if (!__syncthreads_or(any_active)) return;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    const int nkb = (Kdim + TK - 1) / TK;
    for (int kb0 = 0; kb0 < nkb; kb0 += kTcThreads) {
        const int kb = kb0 + tid;
        bool nz = false;
        if (kb < nkb) {
#pragma unroll
            for (int hh = 0; hh < (TN + TILE - 1) / TILE; ++hh) {
                const int ot = (o0 + hh * TILE) / TILE;
                if (o0 + hh * TILE < Odim) nz |= (MODE == 0 ? w.nz[ot * w.nzc + kb] : w.nz[kb * w.nzc + ot]) != 0;
            }
        }
        const unsigned m = __ballot_sync(0xffffffffu, nz);
        if (lane == 0) nzbits[(kb0 >> 5) + warp] = m;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    auto next_block = [&](int k0) {
        int kb = k0 / TK;
        while (kb < nkb) {
            const uint32_t wbits = nzbits[kb >> 5] >> (kb & 31);
            if (wbits) return (kb + __ffs(wbits) - 1) * TK;
            kb = (kb | 31) + 1;
        }
        return Kdim;
    };

    unsigned char* Xp[4] = {smem + kSmemX, smem + kSmemX + kXPlane, smem + kSmemX + 2 * kXPlane, smem + kSmemX + 3 * kXPlane};
    unsigned char* Bp[4] = {smem + kSmemB, smem + kSmemB + kBPlane, smem + kSmemB + 2 * kBPlane, smem + kSmemB + 3 * kBPlane};
    constexpr uint32_t kXLbo = TM * 16, kBLbo = TN * 16, kSbo = 128;
    const uint32_t d_re = tmem_base, d_im = tmem_base + TN;
    uint64_t xd[4], bd[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        xd[q] = umma_desc(smem_u32(Xp[q]), kXLbo, kSbo);
        bd[q] = umma_desc(smem_u32(Bp[q]), kBLbo, kSbo);
    }
    // raw ring: stage = [TM rows of the frame tile | TN rows of the design tile], kRawRow bytes per row
    auto issue = [&](int k0, int stg) {
        unsigned char* raw = smem + RS::ring + stg * RS::stage;
        if (k0 < Kdim) {
#pragma unroll
            for (int u = 0; u < TM * 16 / kTcThreads; ++u) {
                const int id = tid + u * kTcThreads;
                const int r = id >> 4, pc = id & 15;
                const long long f = f0 + r;
                const int k = k0 + pc * 2;
                const bool ok = f < F && k + 1 < Kdim;
                cp_async16(raw + r * kRawRow + pc * 16, ok ? (const void*)(Xin + f * Kdim + k) : (const void*)Xin, ok);
            }
#pragma unroll
            for (int u = 0; u < TN * 16 / kTcThreads; ++u) {
                const int id = tid + u * kTcThreads;
                const int r = id >> 4, pc = id & 15;
                const int o = o0 + r;
                const int k = k0 + pc * 2;
                const bool ok = o < Odim && k + 1 < Kdim;
                cp_async16(raw + (TM + r) * kRawRow + pc * 16, ok ? (const void*)(Bm + (size_t)o * Kdim + k) : (const void*)Bm, ok);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto convert = [&](int stg) {
        const unsigned char* raw = smem + RS::ring + stg * RS::stage;
#pragma unroll
        for (int u = 0; u < TM * (TK / 4) / kTcThreads; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TM, c = it / TM;
            const float4 v0 = *reinterpret_cast<const float4*>(raw + r * kRawRow + c * 32);
            const float4 v1 = *reinterpret_cast<const float4*>(raw + r * kRawRow + c * 32 + 16);
            const uint32_t off = (uint32_t)c * kXLbo + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16;
            split_store(Xp[0], Xp[1], off, v0.x, v0.z, v1.x, v1.z);
            split_store(Xp[2], Xp[3], off, v0.y, v0.w, v1.y, v1.w);
        }
#pragma unroll
        for (int u = 0; u < TN * (TK / 4) / kTcThreads; ++u) {
            const int it = tid + u * kTcThreads;
            const int r = it % TN, c = it / TN;
            const float4 v0 = *reinterpret_cast<const float4*>(raw + (TM + r) * kRawRow + c * 32);
            const float4 v1 = *reinterpret_cast<const float4*>(raw + (TM + r) * kRawRow + c * 32 + 16);
            const uint32_t off = (uint32_t)c * kBLbo + (uint32_t)(r >> 3) * kSbo + (uint32_t)(r & 7) * 16;
            split_store(Bp[0], Bp[1], off, v0.x, v0.z, v1.x, v1.z);
            split_store(Bp[2], Bp[3], off, v0.y, v0.w, v1.y, v1.w);
        }
    };

    // two-stage ring: block i+1 is requested at the top of iteration i (its stage was consumed in iteration i-1), so its
    // loads fly under the conversion and the MMAs of block i
    int k_cur = next_block(0);
    issue(k_cur, 0);
    uint32_t phase = 0;
    int blocks_done = 0;
    while (k_cur < Kdim) {
        const int k_next = next_block(k_cur + TK);
        issue(k_next, (blocks_done + 1) % kRingStages);
        asm volatile("cp.async.wait_group 1;" ::: "memory");   // everything but the request just made
        __syncthreads();                                   // everyone's copies of this block have landed
        if (blocks_done > 0) {                             // the previous block's MMAs still read the planes
            mbar_wait(bar, phase);
            phase ^= 1u;
        }
        convert(blocks_done % kRingStages);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int j = 0; j < TK / 8; ++j) {
                const uint64_t xo = (uint64_t)((2 * j) * kXLbo >> 4), bo = (uint64_t)((2 * j) * kBLbo >> 4);
                const uint32_t acc = (blocks_done > 0 || j > 0) ? 1u : 0u;
                const int xs[4] = {0, 2, 0, 2}, bs[4] = {0, 2, 2, 0};
                const bool ng[4] = {false, MODE == 0, MODE == 1, false};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t dst = (q < 2) ? d_re : d_im;
                    const uint32_t accq = (q == 0 || q == 2) ? acc : 1u;
                    const uint32_t id = ng[q] ? umma_idesc<TN>(true) : umma_idesc<TN>(false);
                    umma_tf32(dst, xd[xs[q]] + xo, bd[bs[q]] + bo, id, accq);
                    umma_tf32(dst, xd[xs[q]] + xo, bd[bs[q] + 1] + bo, id, 1u);
                    umma_tf32(dst, xd[xs[q] + 1] + xo, bd[bs[q]] + bo, id, 1u);
                }
            }
            umma_commit(bar);
        }
        ++blocks_done;
        k_cur = k_next;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (blocks_done > 0) {
        mbar_wait(bar, phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }

    float2* tile = reinterpret_cast<float2*>(smem);            // [TM][TN + 1] complex sums (the planes are free now)
    if (warp < 4 * (TN / 32)) {
        const int fr = (warp & 3) * 32 + lane;
        const int half = warp >> 2;
        uint32_t vr[32], vi[32];
        if (blocks_done > 0) {
            const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(half * 32);
            tmem_ld32(taddr, vr);
            tmem_ld32(taddr + TN, vi);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
            for (int q = 0; q < 32; ++q) vr[q] = vi[q] = 0u;
        }
#pragma unroll
        for (int q = 0; q < 32; ++q) tile[fr * (TN + 1) + half * 32 + q] = make_float2(__uint_as_float(vr[q]), __uint_as_float(vi[q]));
    }
    __syncthreads();
    tc_epilogue<MODE, TN>(w, g, y, F, tile, act_s, f0, o0, Odim, tid);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

__global__ void transpose_c64_kernel(const float2* __restrict__ A, float2* __restrict__ At, int n, int N) {
    __shared__ float2 tile[32][33];
    const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, j = j0 + threadIdx.x;
        tile[r][threadIdx.x] = (i < n && j < N) ? A[(size_t)i * N + j] : make_float2(0.f, 0.f);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int j = j0 + r, i = i0 + threadIdx.x;
        if (i < n && j < N) At[(size_t)j * n + i] = tile[threadIdx.x][r];
    }
}

}  // namespace

bool scamp_tc_fits(int n, int N, long long F) {
    // both reductions must fit the non-zero bitmap, and the frame tiles the grid's y dimension
    return (n + TK - 1) / TK <= kMaxKBlocks && (N + TK - 1) / TK <= kMaxKBlocks && (F + TM - 1) / TM <= 65535;
}

int scamp_tc_prepare(const float2* A, float2* At, int n, int N, cudaStream_t stream) {
    transpose_c64_kernel<<<dim3((N + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, stream>>>(A, At, n, N);
    count_launch();
    return check_cuda(cudaGetLastError(), "transpose_c64_kernel launch");
}

int scamp_tc_gemm(int mode, const ScampWs& w, const Geom& g, const float2* Bm, const float2* y, long long F, cudaStream_t stream) {
    // the residual GEMM has few outputs (n): 32-wide tiles give it twice the CTAs (two per SM overlap staging and MMAs)
    constexpr int TN0 = 32, TN1 = 64;
    // function attributes are per device: set on every launch (cheap), as the other launchers do
    if (int e = check_cuda(cudaFuncSetAttribute(scamp_tc_gemm_kernel<0, TN0>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<TN0>::total),
                           "cudaFuncSetAttribute(scamp_tc<0>)"))
        return e;
    if (int e = check_cuda(cudaFuncSetAttribute(scamp_tc_gemm_kernel<1, TN1>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcSmem<TN1>::total),
                           "cudaFuncSetAttribute(scamp_tc<1>)"))
        return e;
    const int Kdim = mode == 0 ? g.N : g.n;
    if ((Kdim + TK - 1) / TK > kMaxKBlocks) {
        set_error("SCAMP tensor-core GEMM: reduction length %d exceeds %d", Kdim, kMaxKBlocks * TK);
        return AMPSM_ENOFIT;
    }
    const int Odim = mode == 0 ? g.n : g.N;
    const int TNm = mode == 0 ? TN0 : TN1;
    const dim3 grid((Odim + TNm - 1) / TNm, (unsigned)((F + TM - 1) / TM));
    if (mode == 0 && (Kdim & 1) == 0 && reinterpret_cast<uintptr_t>(Bm) % 16 == 0 && !getenv("AMPSM_SCAMP_NORING")) {
        if (int e = check_cuda(cudaFuncSetAttribute(scamp_tc_ring_kernel<0, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                    RingSmem<64>::total), "cudaFuncSetAttribute(scamp_tc_ring)"))
            return e;
        const dim3 grid_r((Odim + 63) / 64, (unsigned)((F + TM - 1) / TM));     // 64-wide tiles: one wave of CTAs at 1024 frames
        scamp_tc_ring_kernel<0, 64><<<grid_r, kTcThreads, RingSmem<64>::total, stream>>>(w, g, Bm, y, F);
    } else if (mode == 0)
        scamp_tc_gemm_kernel<0, TN0><<<grid, kTcThreads, TcSmem<TN0>::total, stream>>>(w, g, Bm, y, F);
    else
        scamp_tc_gemm_kernel<1, TN1><<<grid, kTcThreads, TcSmem<TN1>::total, stream>>>(w, g, Bm, y, F);
    count_launch();
    return check_cuda(cudaGetLastError(), "scamp_tc_gemm_kernel launch");
}

}  // namespace ampsm
