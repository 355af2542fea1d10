// SCAMP on the reference's own coupled design matrix, applied from its taps (channel.py:76-96):
//   A = sum_l kron(eye(Lout, Lin, -l) * sqrt(W), h_l)   =>   block (r, c) of A = T_{r-c} for 0 <= r - c < Lh,  T_l = sqrt(W_l) h_l (Nr x Nt)
// so both mat-vecs of an iteration (scamp.py:48,56) are ONE dense complex GEMM each over the rows (frame f, column block c):
//   residual : P[(f,c)][l Nr + i] = sum_j  xh[f][c Nt + j] T_l[i][j]            K = Nt,     N = Lh Nr   then  S[f][r Nr + i] = sum_l P[(f,r-l)][l Nr + i]
//   estimate : Q[(f,c)][j]        = sum_{l,i} zs[f][(c+l) Nr + i] conj(T_l[i][j])  K = Lh Nr,  N = Nt      (tap l of row (f,c) reads row block c + l of zs[f])
// The dense A (142 MB at BASELINE config 4) is never read: the taps are 393 KB and stay in L2.
//
// One kernel, two modes, warp-specialised, everything asynchronous:
//   warp 8 (one lane)  : TMA producer.  The raw complex64 operand tile (128 rows x 16 reduction elements = 128-byte rows,
//                        SWIZZLE_128B) arrives by cp.async.bulk.tensor (2-D map over the xh rows / 3-D map
//                        (element in block, row block, frame) over zs, tap l = a shift of the block coordinate); the design operand arrives pre-split (hi / lo, re / im planes in
//                        the UMMA canonical layout, written once per call) by one cp.async.bulk per stage.
//   warps 4-7          : converters, thread = row: raw tile -> {re,im} x {hi,lo} float32 planes (3xTF32 split: hi = value with
//                        the 13 low mantissa bits cleared, lo = value - hi, exact) in the canonical K-major no-swizzle layout.
//   warp 9 (one lane)  : MMA issuer: 12 tcgen05.mma kind::tf32 per K = 8 step (4 real products x 3 split terms), float32
//                        accumulators in TMEM, tcgen05.commit releases the plane stage / hands the accumulator to the epilogue.
//   warps 0-3          : epilogue, thread = TMEM lane = row.  Estimate mode: the accumulator of one 128-output chunk is drained
//                        (double-buffered in TMEM, so the next chunk's MMAs run underneath) through a per-warp transpose buffer
//                        and applied as xmap = xh + tau S with whole-line global accesses.  Residual mode: P goes to shared
//                        memory and the CTA (whole frames per tile) forms the overlap-add over the Lh taps and the fused update
//                        z = y - S + b z, zs = z / phi.
//                        With FUSED the section denoiser (scamp.py:61-68), the new estimate, psi (scamp.py:59) and the allclose test
//                        (scamp.py:105) run in this epilogue as well: lanes = antennas, segmented warp reductions per section.
// Rings between the roles (mbarriers): raw tiles 3-6 deep (HBM / L2 latency), split planes 2 deep, design planes 2-4 deep;
// the tile of a CTA is FR = floor(128 / Lin) whole frames.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "blockops.cuh"
#include "fastops.cuh"
#include "kernels.h"
#include "scamp_ws.cuh"

namespace ampsm {

namespace {

constexpr int TM = 128;            // rows per tile (UMMA M)
constexpr int KS = 16;             // complex reduction elements per stage: one 128-byte swizzled row, two K = 8 MMA steps
constexpr int kXStages = 2;        // split-plane ring (X operand)
constexpr int kMaxRaw = 6, kMaxB = 4;
constexpr int kEpiWarps = 8, kCvtWarps = 4;          // two epilogue warps per TMEM lane quarter (the residual mode drains with the first four)
constexpr int kThreads = (kEpiWarps + kCvtWarps + 2) * 32;     // + producer warp + MMA warp
constexpr int kRawStage = TM * KS * 8;                         // 16 KiB
constexpr int kXPlane = TM * KS * 4;                           // 8 KiB: [4 K chunks][128 rows][4 floats]
constexpr int kXStage = 4 * kXPlane;                           // 32 KiB
constexpr int kEpiCols = 32;                                   // accumulator columns drained per tcgen05.ld
constexpr int kTbufRows = 16, kTbufStride = 65;                // per-warp transpose buffer: 16 rows x (64 + 1) complex sums
constexpr int kSmemMax = 227 * 1024;

struct StSmem {          // byte offsets into dynamic shared memory (1024-byte aligned base)
    int raw, xpl, bpl, epi, meta, bars, total;
    int nraw, nb;        // ring depths
};
__host__ __device__ inline StSmem st_smem(int brows, int mode, int force_raw = 0, int force_b = 0) {
    StSmem s;
    const int bstage = 4 * KS * brows * 4;                     // B stage: 4 planes x [4 K chunks][brows][4 floats]
    const int epi_bytes = mode == 1 ? kEpiWarps * kTbufRows * kTbufStride * 8 : 0;
    const int meta_bytes = TM * 24;                            // per row: global offset (8), frame (4), tau (4), psi (4), energy (4)
    const int fixed = kXStages * kXStage + epi_bytes + meta_bytes + 256;
    // deepest rings that fit: raw tiles first (they come from HBM), then the design planes (L2)
    s.nraw = 3;
    s.nb = 2;
    for (bool grew = true; grew;) {
        grew = false;
        if (s.nraw < kMaxRaw && fixed + (s.nraw + 1) * kRawStage + s.nb * bstage <= kSmemMax) { ++s.nraw; grew = true; }
        if (s.nb < kMaxB && fixed + s.nraw * kRawStage + (s.nb + 1) * bstage <= kSmemMax) { ++s.nb; grew = true; }
    }
    if (force_raw >= 2 && force_b >= 1 && force_raw <= kMaxRaw && force_b <= kMaxB &&
        fixed + force_raw * kRawStage + force_b * bstage <= kSmemMax) {       // experiments: AMPSM_ST_RINGS="raw,b"
        s.nraw = force_raw;
        s.nb = force_b;
    }
    s.raw = 0;
    s.xpl = s.raw + s.nraw * kRawStage;
    s.bpl = s.xpl + kXStages * kXStage;
    s.epi = s.bpl + s.nb * bstage;
    s.meta = s.epi + epi_bytes;
    s.bars = s.meta + meta_bytes;
    if (mode == 0) {                                           // P staging [128][brows + 1] float2 aliases the rings
        const int pbytes = TM * (brows + 1) * 8;
        if (pbytes > s.meta) {
            s.meta = (pbytes + 15) & ~15;
            s.bars = s.meta + meta_bytes;
        }
    }
    s.total = s.bars + 256;
    return s;
}

__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// D = F32, A = B = TF32, both K-major, M = 128, N = n, optional negate-A
__device__ __forceinline__ uint32_t umma_idesc(int n, bool neg_a) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                 : "memory");
}

struct StArgs {
    ScampWs w;
    Geom g;
    DevAlphabet al;
    const float2* y;
    const unsigned char* bplanes;      // pre-split design operand of this mode: [chunk][stage][4 planes][4 K chunks][brows][4 floats]
    long long F;
    int Lh, FR;                        // taps, frames per tile
    int kpb;                           // estimate mode: K stages per tap (row block of zs padded to a multiple of KS)
    int brows;                         // N of the MMAs: outputs per chunk (padded to 32)
    int chunks, nks;                   // output chunks per tile, K stages per chunk
    int zs_stride;                     // complex elements between frames of Zs (padded to (Lin + Lh - 1) Nr)
    const float* W;                    // [Lout][Lin] base matrix (residual mode computes the block scalars of its frames)
    float sigma2;
    const float* sigma2_pf;
    int t;                             // iteration index (the fused estimate mode retires its frames itself)
    int ring_raw, ring_b;              // forced ring depths (0: automatic)
    int dbg;                           // AMPSM_ST_DEBUG bits (timing experiments): 2 skip the MMAs, 4 skip the conversion, 8 skip the Xh loads, 16 skip the epilogue stores, 32 skip the denoiser math
};

// x - shift for x = q.re s.re + q.im s.im, the symbol given as float32 value + float32 residual of its float64 value: the
// products are split exactly by FMA (hi + lo), the sum by a two-sum, so the difference to the shift carries the accuracy of
// the reference's float64 evaluation (bamp.py:69 / scamp.py:64) without leaving the FP32 pipe (float64 conversions run on the
// XU pipe at a sixteenth of the FP32 rate and bound the first version of this epilogue)
__device__ __forceinline__ float exponent_diff(float qr, float qi, float sr, float si, float srl, float sil, float shift) {
    const float h1 = __fmul_rn(qr, sr), l1 = fmaf(qr, sr, -h1);            // intrinsics: never contracted into other FMAs
    const float h2 = __fmul_rn(qi, si), l2 = fmaf(qi, si, -h2);
    const float sum = __fadd_rn(h1, h2), t = __fsub_rn(sum, h1);
    const float err = __fadd_rn(__fsub_rn(h1, __fsub_rn(sum, t)), __fsub_rn(h2, t));
    const float low = fmaf(qr, srl, fmaf(qi, sil, __fadd_rn(__fadd_rn(l1, l2), err)));
    return __fadd_rn(__fsub_rn(sum, shift), low);
}

// reciprocal to one ulp without a slow path (the slow-path call of __frcp_rn sits behind a branch, which costs the epilogue its
// lock step): MUFU.RCP + one Newton step
__device__ __forceinline__ float rcp1(float x) {
    const float r = fast_rcp(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}

// MODE 0: residual (X = xh rows, map 2-D);  MODE 1: estimate (X = zs row blocks, map 3-D).  FUSED (estimate only): section
// denoiser + psi + exit test in the epilogue (float32 exp, per-section shift; M in {8, 16, 32, 64}).
// MSEC: section size of the fused denoiser (8, 16, 32, 64), 0 = not fused.  EXACT: every symbol is one of {0, +-1, +-j} (the reference's
// OOK / BPSK / QPSK tables, config.py:86-95): the exponent q.re s.re + q.im s.im is then ONE exact float32 product, so the plain float32
// difference to the shift is as accurate as the reference's float64 evaluation and the compensated sum is not needed.  The
// reference's QPSK table {1, j, -1, -j} (config.py:93) is special-cased further: the four exponents are +-q.re, +-q.im, their
// maximum is max(|q.re|, |q.im|), the symbol-weighted sums are differences of two exponentials.
template <int MODE, int MSEC, int ALPH>      // ALPH: 0 compensated exponents, 1 exact products, 2 the four axis symbols {1, j, -1, -j}
__global__ void __launch_bounds__(kThreads, 1) scamp_st_kernel(const __grid_constant__ StArgs a, const __grid_constant__ CUtensorMap xmap_desc) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ int any_active_s;
    const ScampWs& w = a.w;
    const Geom& g = a.g;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rows = a.FR * g.Lin;                                  // valid rows of a tile (<= 128)
    const long long f0 = (long long)blockIdx.x * a.FR;
    const StSmem L = st_smem(a.brows, MODE, a.ring_raw, a.ring_b);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.bars);
    uint64_t *raw_full = bars, *raw_empty = bars + kMaxRaw, *xp_full = bars + 2 * kMaxRaw, *xp_empty = xp_full + kXStages,
             *b_full = xp_empty + kXStages, *b_empty = b_full + kMaxB, *acc_full = b_empty + kMaxB, *acc_empty = acc_full + 2;
    // per-row metadata of the tile: frame (or -1), offset of the row's column block in Xh / Xmap, tau, old psi, energy
    long long* row_off = reinterpret_cast<long long*>(smem + L.meta);
    int* row_f = reinterpret_cast<int*>(smem + L.meta + TM * 8);
    float* row_tau = reinterpret_cast<float*>(smem + L.meta + TM * 12);
    float* row_psi = reinterpret_cast<float*>(smem + L.meta + TM * 16);
    float* row_e = reinterpret_cast<float*>(smem + L.meta + TM * 20);

    // tiles whose frames have all met the exit test are skipped before anything is allocated
    if (tid == 0) any_active_s = 0;
    __syncthreads();
    if (tid < TM) {
        const int r = tid;
        const long long f = f0 + r / g.Lin;
        const int c = r % g.Lin;
        const bool ok = r < rows && f < a.F && w.active[f];
        row_f[r] = ok ? (int)(f - f0) : -1;
        row_off[r] = ok ? f * (long long)g.N + (long long)c * g.Nt : 0;
        if (MODE == 1) {
            row_tau[r] = ok ? w.tau[f * g.Lin + c] : 0.f;
            row_psi[r] = ok ? w.psi[f * g.Lin + c] : 0.f;
            row_e[r] = 0.f;
        }
        if (ok) any_active_s = 1;
    }
    __syncthreads();
    if (!any_active_s) return;

    const int tmem_cols_needed = (MODE == 1 ? 2 : 1) * 2 * a.brows;
    const uint32_t tmem_cols = tmem_cols_needed <= 32 ? 32 : tmem_cols_needed <= 64 ? 64 : tmem_cols_needed <= 128 ? 128 : tmem_cols_needed <= 256 ? 256 : 512;
    if (warp == kEpiWarps + kCvtWarps + 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int s = 0; s < kMaxRaw; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], kCvtWarps * 32);
        }
        for (int s = 0; s < kXStages; ++s) {
            mbar_init(&xp_full[s], kCvtWarps * 32);
            mbar_init(&xp_empty[s], 1);
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], kEpiWarps * 32);      // estimate mode: both warp sets read every accumulator buffer
        }
        for (int s = 0; s < kMaxB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        fence_mbar_init();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const int bstage_bytes = 4 * KS * a.brows * 4;
    const int total_it = a.chunks * a.nks;
    const int NR = L.nraw, NB = L.nb;

    if (warp == kEpiWarps + kCvtWarps) {
        // ===================================================== TMA producer
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap_desc) : "memory");
            for (int it = 0; it < total_it; ++it) {
                const int ks = it % a.nks;
                const int rs = it % NR, ru = it / NR;
                mbar_wait(&raw_empty[rs], (ru & 1) ^ 1);
                mbar_expect_tx(&raw_full[rs], (uint32_t)(rows * KS * 8));
                if (MODE == 0) tma_load_2d(smem + L.raw + rs * kRawStage, &xmap_desc, ks * KS * 2, (int)(f0 * g.Lin), &raw_full[rs]);
                else tma_load_3d(smem + L.raw + rs * kRawStage, &xmap_desc, (ks % a.kpb) * KS * 2, ks / a.kpb, (int)f0, &raw_full[rs]);
                const int bs = it % NB, bu = it / NB;
                mbar_wait(&b_empty[bs], (bu & 1) ^ 1);
                mbar_expect_tx(&b_full[bs], (uint32_t)bstage_bytes);
                tma_load_1d(smem + L.bpl + bs * bstage_bytes, a.bplanes + (size_t)it * bstage_bytes, (uint32_t)bstage_bytes, &b_full[bs]);
            }
        }
    } else if (warp == kEpiWarps + kCvtWarps + 1) {
        // ===================================================== MMA issuer
        if (lane == 0) {
            const uint32_t id_pos = umma_idesc(a.brows, false), id_neg = umma_idesc(a.brows, true);
            const uint32_t x_lbo = TM * 16, b_lbo = (uint32_t)a.brows * 16, sbo = 128;
            const uint32_t bplane = 4 * (uint32_t)a.brows * 16;          // bytes of one B plane of a stage
            const uint64_t xdesc0 = umma_desc(smem_u32(smem + L.xpl), x_lbo, sbo), bdesc0 = umma_desc(smem_u32(smem + L.bpl), b_lbo, sbo);
            for (int ch = 0; ch < a.chunks; ++ch) {
                const int buf = MODE == 1 ? (ch & 1) : 0, v = MODE == 1 ? (ch >> 1) : ch;
                mbar_wait(&acc_empty[buf], (v & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_re = tmem_base + (uint32_t)(buf * 2 * a.brows), d_im = d_re + (uint32_t)a.brows;
                for (int ks = 0; ks < a.nks; ++ks) {
                    const int it = ch * a.nks + ks;
                    const int ps = it % kXStages, pu = it / kXStages;
                    const int bs = it % NB, bu = it / NB;
                    mbar_wait(&xp_full[ps], pu & 1);
                    mbar_wait(&b_full[bs], bu & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    // descriptors differ only in their start-address field (bits 0-13, units of 16 bytes): base + offset
                    const uint64_t xd0 = xdesc0 + (uint64_t)((ps * kXStage) >> 4), bd0 = bdesc0 + (uint64_t)((bs * bstage_bytes) >> 4);
#pragma unroll
                    for (int j = 0; j < ((a.dbg & 2) ? 0 : KS / 8); ++j) {
                        // planes: 0 re_hi, 1 re_lo, 2 im_hi, 3 im_lo;  re = XrBr - XiBi, im = XrBi + XiBr  (the estimate mode's conjugate
                        // lives in its pre-split B planes)
                        uint64_t xd[4], bd[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            xd[q] = xd0 + (uint64_t)((q * kXPlane + 2 * j * x_lbo) >> 4);
                            bd[q] = bd0 + (uint64_t)((q * bplane + 2 * j * b_lbo) >> 4);
                        }
                        const uint32_t acc = (ks > 0 || j > 0) ? 1u : 0u;
                        const int xsel[4] = {0, 2, 0, 2}, bsel[4] = {0, 2, 2, 0};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t dst = (q < 2) ? d_re : d_im;
                            const uint32_t accq = (q == 0 || q == 2) ? acc : 1u;
                            const uint32_t id = (q == 1) ? id_neg : id_pos;
                            umma_tf32(dst, xd[xsel[q]], bd[bsel[q]], id, accq);            // hi * hi
                            umma_tf32(dst, xd[xsel[q]], bd[bsel[q] + 1], id, 1u);          // hi * lo
                            umma_tf32(dst, xd[xsel[q] + 1], bd[bsel[q]], id, 1u);          // lo * hi
                        }
                    }
                    umma_commit(&xp_empty[ps]);                // both operand stages are free once these MMAs have read them
                    umma_commit(&b_empty[bs]);
                }
                umma_commit(&acc_full[buf]);                   // ... and the accumulator is complete
            }
        }
    } else if (warp >= kEpiWarps) {
        // ===================================================== converters: thread = row
        const int r = tid - kEpiWarps * 32;
        const bool live = r < rows;
        for (int it = 0; it < total_it; ++it) {
            const int rs = it % NR, ru = it / NR;
            const int ps = it % kXStages, pu = it / kXStages;
            mbar_wait(&raw_full[rs], ru & 1);
            float4 v[8];
            const unsigned char* raw = smem + L.raw + rs * kRawStage + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
                v[j] = live ? *reinterpret_cast<const float4*>(raw + ((j ^ (r & 7)) << 4)) : make_float4(0.f, 0.f, 0.f, 0.f);   // SWIZZLE_128B
            mbar_arrive(&raw_empty[rs]);
            mbar_wait(&xp_empty[ps], (pu & 1) ^ 1);
            unsigned char* xp = smem + L.xpl + ps * kXStage + r * 16;
#pragma unroll
            for (int c = 0; c < ((a.dbg & 4) ? 0 : 4); ++c) {                      // K chunk c = elements 4c .. 4c+3 = raw chunks 2c, 2c+1
                const float4 p = v[2 * c], q = v[2 * c + 1];
                const float re[4] = {p.x, p.z, q.x, q.z}, im[4] = {p.y, p.w, q.y, q.w};
                float rh[4], rl[4], ih[4], il[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    rh[e] = __uint_as_float(__float_as_uint(re[e]) & 0xffffe000u);
                    rl[e] = re[e] - rh[e];
                    ih[e] = __uint_as_float(__float_as_uint(im[e]) & 0xffffe000u);
                    il[e] = im[e] - ih[e];
                }
                unsigned char* dst = xp + c * (TM * 16);
                *reinterpret_cast<float4*>(dst) = make_float4(rh[0], rh[1], rh[2], rh[3]);
                *reinterpret_cast<float4*>(dst + kXPlane) = make_float4(rl[0], rl[1], rl[2], rl[3]);
                *reinterpret_cast<float4*>(dst + 2 * kXPlane) = make_float4(ih[0], ih[1], ih[2], ih[3]);
                *reinterpret_cast<float4*>(dst + 3 * kXPlane) = make_float4(il[0], il[1], il[2], il[3]);
            }
            fence_proxy_async();                               // generic-proxy stores -> visible to the tensor core
            mbar_arrive(&xp_full[ps]);
        }
    } else {
        // ===================================================== epilogue warps 0-3: thread = TMEM lane = row
        const int quarter = warp & 3, eset = warp >> 2;        // TMEM lanes 32 quarter ..; set 0 takes rows 0-15 of the quarter, set 1 rows 16-31
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        if (MODE == 1) {
            // per-warp transpose buffer: 16 rows x 64 columns (+1 pad) complex sums
            float2* tbuf = reinterpret_cast<float2*>(smem + L.epi) + warp * kTbufRows * kTbufStride;
            constexpr bool FUSED = MSEC > 0;
            constexpr int M = MSEC, segw = MSEC < 32 ? MSEC : 32;
            const int K = a.al.K;
            const int pairs = (a.brows + 63) / 64;             // 64-column pairs of 32-column groups per chunk
            auto prefetch_chunk = [&](int ch) {                // this thread's row: the chunk's segment of Xh -> L2
                    const int cols = min(a.brows, g.Nt - ch * a.brows);
                    if (tid < TM && ch < a.chunks && row_f[tid] >= 0 && cols > 0 && ((cols * 8) & 15) == 0) {
                        const float2* src = w.Xh + row_off[tid] + (long long)ch * a.brows;
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(cols * 8) : "memory");
                    }
            };
            prefetch_chunk(0);
            for (int ch = 0; ch < a.chunks; ++ch) {
                const int buf = ch & 1, v = ch >> 1;
                prefetch_chunk(ch + 1);
                mbar_wait(&acc_full[buf], v & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t col0 = (uint32_t)(buf * 2 * a.brows);
                for (int pr = 0; pr < pairs; ++pr) {
                    const bool haveB = pr * 64 + 32 < a.brows;
                    const int oA = ch * a.brows + pr * 64 + lane, oB = oA + 32;
                    const bool okA = oA < g.Nt, okB = haveB && oB < g.Nt;
                    const bool ldA = okA && !(a.dbg & 8), ldB = okB && !(a.dbg & 8), stA = okA && !(a.dbg & 16), stB = okB && !(a.dbg & 16);
                    // the 32 rows of this warp in two passes of 16: the lanes of the pass park their row (64 complex sums) in the
                    // transpose buffer, then the warp walks the 16 rows with lane = antenna (columns oA = lane, oB = lane + 32)
                    {
                        const int pass = eset;
                        {
                            uint32_t vr[32], vi[32];
                            tmem_ld32(lane_addr + col0 + pr * 64, vr);
                            tmem_ld32(lane_addr + col0 + a.brows + pr * 64, vi);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            __syncwarp();
                            if ((lane >> 4) == pass) {
#pragma unroll
                                for (int q = 0; q < 32; ++q)
                                    tbuf[(lane & 15) * kTbufStride + q] = make_float2(__uint_as_float(vr[q]), __uint_as_float(vi[q]));
                            }
                            if (haveB) {
                                tmem_ld32(lane_addr + col0 + pr * 64 + 32, vr);
                                tmem_ld32(lane_addr + col0 + a.brows + pr * 64 + 32, vi);
                                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                                if ((lane >> 4) == pass) {
#pragma unroll
                                    for (int q = 0; q < 32; ++q)
                                        tbuf[(lane & 15) * kTbufStride + 32 + q] = make_float2(__uint_as_float(vr[q]), __uint_as_float(vi[q]));
                                }
                            }
                        }
                        if (pr == pairs - 1) {                 // last read of this accumulator buffer: hand it back to the MMA warp
                            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                            mbar_arrive(&acc_empty[buf]);
                        }
                        __syncwarp();
                        const int rbase = quarter * 32 + pass * 16;
                        // the previous estimates of four rows at a time, requested one block ahead (their lines were pulled into L2
                        // by the bulk prefetch issued one chunk earlier)
                        float2 pa[4], pb[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            pa[u] = pb[u] = make_float2(0.f, 0.f);
                            if (row_f[rbase + u] >= 0) {
                                if (ldA) pa[u] = w.Xh[row_off[rbase + u] + oA];
                                if (ldB) pb[u] = w.Xh[row_off[rbase + u] + oB];
                            }
                        }
#pragma unroll 1
                        for (int r4 = 0; r4 < kTbufRows; r4 += 4) {
                            float2 ca[4], cb[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                ca[u] = pa[u];
                                cb[u] = pb[u];
                            }
                            if (r4 + 4 < kTbufRows) {
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const int rn = rbase + r4 + 4 + u;
                                    if (row_f[rn] >= 0) {
                                        if (ldA) pa[u] = w.Xh[row_off[rn] + oA];
                                        if (ldB) pb[u] = w.Xh[row_off[rn] + oB];
                                    }
                                }
                            }
                            // rows of frames that met the exit test are frozen: a block of four rows belongs to one frame when
                            // 4 | Lin, so whole blocks drop out as the batch converges
                            if (row_f[rbase + r4] < 0 && row_f[rbase + r4 + 1] < 0 && row_f[rbase + r4 + 2] < 0 && row_f[rbase + r4 + 3] < 0) continue;
                            // four rows in lock step, phase by phase: their dependent chains (shuffle reductions, exponent sums,
                            // MUFU) interleave, which is what keeps the two epilogue warps of a scheduler busy
                            bool live[4];
                            long long at[4];
                            float tau[4];
                            float2 ma[4], mb[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) {
                                const int row = rbase + r4 + u;
                                live[u] = row_f[row] >= 0;
                                at[u] = row_off[row];
                                tau[u] = row_tau[row];
                                const float2 sa = tbuf[(r4 + u) * kTbufStride + lane];
                                const float2 sb = haveB ? tbuf[(r4 + u) * kTbufStride + 32 + lane] : make_float2(0.f, 0.f);
                                ma[u] = make_float2(fmaf(tau[u], sa.x, ca[u].x), fmaf(tau[u], sa.y, ca[u].y));     // scamp.py:56
                                mb[u] = make_float2(fmaf(tau[u], sb.x, cb[u].x), fmaf(tau[u], sb.y, cb[u].y));
                                if (live[u] && stA) w.Xmap[at[u] + oA] = ma[u];
                                if (live[u] && stB) w.Xmap[at[u] + oB] = mb[u];
                            }
                            if (FUSED && !(a.dbg & 32)) {
                                // section-wise posterior mean (scamp.py:61-68): s / (tau / 2) in complex64, exponents as float64 products,
                                // float32 ex2 of the difference to the SECTION maximum, lanes = antennas
                                float qar[4], qai[4], qbr[4], qbi[4], la[4], lb[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float rt = rcp1(tau[u] / 2.0f);
                                    qar[u] = __fmul_rn(ma[u].x, rt);
                                    qai[u] = __fmul_rn(ma[u].y, rt);
                                    qbr[u] = __fmul_rn(mb[u].x, rt);
                                    qbi[u] = __fmul_rn(mb[u].y, rt);
                                    la[u] = lb[u] = -INFINITY;
                                }
                                if (ALPH == 2) {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        la[u] = fmaxf(fabsf(qar[u]), fabsf(qai[u]));
                                        lb[u] = fmaxf(fabsf(qbr[u]), fabsf(qbi[u]));
                                    }
                                } else {
                                    for (int k = 0; k < K; ++k) {
                                        const float sr = a.al.ref[k], si = a.al.imf[k];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            la[u] = fmaxf(la[u], fmaf(qar[u], sr, qai[u] * si));
                                            lb[u] = fmaxf(lb[u], fmaf(qbr[u], sr, qbi[u] * si));
                                        }
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (!okA) la[u] = -INFINITY;
                                    if (!okB) lb[u] = -INFINITY;
                                    if (M == 64) la[u] = lb[u] = fmaxf(la[u], lb[u]);
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    if (o < segw) {
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            la[u] = fmaxf(la[u], __shfl_xor_sync(0xffffffffu, la[u], o));
                                            if (M != 64) lb[u] = fmaxf(lb[u], __shfl_xor_sync(0xffffffffu, lb[u], o));
                                        }
                                    }
                                }
                                float za[4], zb[4], nar[4], nai[4], nbr[4], nbi[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (M == 64) lb[u] = la[u];
                                    za[u] = zb[u] = nar[u] = nai[u] = nbr[u] = nbi[u] = 0.f;
                                }
                                if (ALPH == 2) {
                                    constexpr float L2E = 1.4426950408889634f;
#pragma unroll
                                    for (int u = 0; u < 4; ++u) {
                                        const float a1 = fast_ex2((qar[u] - la[u]) * L2E), a2 = fast_ex2((-qar[u] - la[u]) * L2E);
                                        const float a3 = fast_ex2((qai[u] - la[u]) * L2E), a4 = fast_ex2((-qai[u] - la[u]) * L2E);
                                        const float b1 = fast_ex2((qbr[u] - lb[u]) * L2E), b2 = fast_ex2((-qbr[u] - lb[u]) * L2E);
                                        const float b3 = fast_ex2((qbi[u] - lb[u]) * L2E), b4 = fast_ex2((-qbi[u] - lb[u]) * L2E);
                                        za[u] = (a1 + a2) + (a3 + a4);
                                        nar[u] = a1 - a2;
                                        nai[u] = a3 - a4;
                                        zb[u] = (b1 + b2) + (b3 + b4);
                                        nbr[u] = b1 - b2;
                                        nbi[u] = b3 - b4;
                                    }
                                } else if (ALPH == 1) {
                                    for (int k = 0; k < K; ++k) {
                                        const float sr = a.al.ref[k], si = a.al.imf[k];
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            const float ea = fast_ex2((fmaf(qar[u], sr, qai[u] * si) - la[u]) * 1.4426950408889634f);
                                            const float eb = fast_ex2((fmaf(qbr[u], sr, qbi[u] * si) - lb[u]) * 1.4426950408889634f);
                                            za[u] += ea;
                                            nar[u] = fmaf(sr, ea, nar[u]);
                                            nai[u] = fmaf(si, ea, nai[u]);
                                            zb[u] += eb;
                                            nbr[u] = fmaf(sr, eb, nbr[u]);
                                            nbi[u] = fmaf(si, eb, nbi[u]);
                                        }
                                    }
                                } else {
                                    for (int k = 0; k < K; ++k) {
                                        const float sr = a.al.ref[k], si = a.al.imf[k];
                                        const float srl = a.al.rel[k], sil = a.al.iml[k];          // float32 residuals of the float64 symbols
    #pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            const float ea = fast_ex2(exponent_diff(qar[u], qai[u], sr, si, srl, sil, la[u]) * 1.4426950408889634f);
                                            const float eb = fast_ex2(exponent_diff(qbr[u], qbi[u], sr, si, srl, sil, lb[u]) * 1.4426950408889634f);
                                            za[u] += ea;
                                            nar[u] = fmaf(sr, ea, nar[u]);
                                            nai[u] = fmaf(si, ea, nai[u]);
                                            zb[u] += eb;
                                            nbr[u] = fmaf(sr, eb, nbr[u]);
                                            nbi[u] = fmaf(si, eb, nbi[u]);
                                        }
                                    }
                                }
                                // section sums of Z; the energy of the normalised estimates is (sum |n|^2) / Z^2 per section
                                float en[4];
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (!okA) za[u] = nar[u] = nai[u] = 0.f;
                                    if (!okB) zb[u] = nbr[u] = nbi[u] = 0.f;
                                    if (M == 64) za[u] = zb[u] = za[u] + zb[u];
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
                                    if (o < segw) {
#pragma unroll
                                        for (int u = 0; u < 4; ++u) {
                                            za[u] += __shfl_xor_sync(0xffffffffu, za[u], o);
                                            if (M != 64) zb[u] += __shfl_xor_sync(0xffffffffu, zb[u], o);
                                        }
                                    }
                                }
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    if (M == 64) zb[u] = za[u];
                                    const float rza = rcp1(za[u]), rzb = rcp1(zb[u]);
                                    const float2 va = make_float2(nar[u] * rza, nai[u] * rza), vb = make_float2(nbr[u] * rzb, nbi[u] * rzb);
                                    if (live[u] && stA) w.Xh[at[u] + oA] = va;
                                    if (live[u] && stB) w.Xh[at[u] + oB] = vb;
                                    en[u] = (okA ? fmaf(va.x, va.x, va.y * va.y) : 0.f) + (okB ? fmaf(vb.x, vb.x, vb.y * vb.y) : 0.f);
                                }
#pragma unroll
                                for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                                    for (int u = 0; u < 4; ++u) en[u] += __shfl_xor_sync(0xffffffffu, en[u], o);
                                }
                                if (lane < 4) {
                                    const float e = lane == 0 ? en[0] : lane == 1 ? en[1] : lane == 2 ? en[2] : en[3];
                                    const bool lv = lane == 0 ? live[0] : lane == 1 ? live[1] : lane == 2 ? live[2] : live[3];
                                    if (lv) row_e[rbase + r4 + lane] += e;
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
            }
            if (MSEC > 0) {
                // psi of the column block and its allclose test (scamp.py:59,105): lane rr finishes row rr of this warp
                __syncwarp();
                const int row = quarter * 32 + eset * 16 + (lane & 15);
                if (lane < 16 && row_f[row] >= 0) {
                    const long long f = f0 + row_f[row];
                    const int c = row % g.Lin;
                    const float pn = 1.0f - row_e[row] / (float)g.Na;
                    const float po = row_psi[row];
                    if (!(fabsf(pn - po) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, po))))) w.notclose[f] = 1;
                    w.psi[f * g.Lin + c] = pn;
                }
            }
        } else if (warp < 4) {
            // P -> shared memory [row][brows + 1] (the rings are idle once acc_full has fired: every MMA has completed)
            const int r = tid;
            mbar_wait(&acc_full[0], 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float2* P = reinterpret_cast<float2*>(smem);
            const int pstride = a.brows + 1;
            for (int gq = 0; gq < a.brows / kEpiCols; ++gq) {
                uint32_t vr[32], vi[32];
                tmem_ld32(lane_addr + gq * kEpiCols, vr);
                tmem_ld32(lane_addr + a.brows + gq * kEpiCols, vi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int q = 0; q < 32; ++q) P[r * pstride + gq * kEpiCols + q] = make_float2(__uint_as_float(vr[q]), __uint_as_float(vi[q]));
            }
        } else if (MODE == 0) {
            // warps 4-7, idle in this mode: the block scalars of the tile's frames under the main loop (scamp.py:45-52):
            // gamma = W psi / Lc, b = gamma / phi_old, phi = sigma2 + gamma, tau = L / (W^T (1 / phi)) / Mr
            const int j = tid - 4 * 32;
            const int Lr = g.Lout, Lc = g.Lin;
            for (int e = j; e < a.FR * Lr; e += 128) {
                const int fl = e / Lr, r = e % Lr;
                if (row_f[fl * Lc] < 0) continue;
                const long long f = f0 + fl;
                float acc = 0.f;
                for (int c = 0; c < Lc; ++c) acc = fmaf(a.W[r * Lc + c], w.psi[f * Lc + c], acc);
                const float gma = acc / (float)Lc;
                const float s2 = a.sigma2_pf ? a.sigma2_pf[f] : a.sigma2;
                w.b[f * Lr + r] = gma / w.phi[f * Lr + r];          // old phi (inf on the first pass -> 0)
                w.phi[f * Lr + r] = s2 + gma;
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");          // the four warps of this role: phi is complete
            for (int e = j; e < a.FR * Lc; e += 128) {
                const int fl = e / Lc, c = e % Lc;
                if (row_f[fl * Lc] < 0) continue;
                const long long f = f0 + fl;
                float acc = 0.f;
                for (int r = 0; r < Lr; ++r) acc = fmaf(a.W[r * Lc + c], __frcp_rn(w.phi[f * Lr + r]), acc);
                w.tau[f * Lc + c] = (float)g.L / acc / (float)g.Nr;      // L = Na*Lin, Mr = Nr (scamp.py:52)
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (MODE == 0) {
        // overlap-add over the taps + fused update (scamp.py:48): all threads, consecutive threads on consecutive outputs
        const float2* P = reinterpret_cast<const float2*>(smem);
        const int pstride = a.brows + 1;
        const int per_frame = g.Lout * g.Nr;                  // = n
        // four outputs per thread at a time: every global load of a batch is issued before the first use
        constexpr int kB = 4;
        const int total = a.FR * per_frame;
        for (int e0 = tid; e0 < total; e0 += kThreads * kB) {
            float2 yv[kB], zo[kB];
            float bb[kB], ph[kB];
            bool ok[kB];
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                const int e = e0 + u * kThreads;
                const int fl = e / per_frame, o = e % per_frame;
                const long long f = f0 + fl;
                ok[u] = e < total && f < a.F && row_f[fl * g.Lin] >= 0;
                yv[u] = zo[u] = make_float2(0.f, 0.f);
                bb[u] = 0.f;
                ph[u] = 1.f;
                if (ok[u]) {
                    const int rb = o / g.Nr;
                    yv[u] = a.y[f * g.n + o];
                    zo[u] = w.Z[f * g.n + o];
                    bb[u] = w.b[f * g.Lout + rb];
                    ph[u] = w.phi[f * g.Lout + rb];
                }
            }
#pragma unroll
            for (int u = 0; u < kB; ++u) {
                if (!ok[u]) continue;
                const int e = e0 + u * kThreads;
                const int fl = e / per_frame, o = e % per_frame;
                const long long f = f0 + fl;
                const int rb = o / g.Nr, i = o % g.Nr;
                float sx = 0.f, sy = 0.f;
                for (int l = 0; l < a.Lh; ++l) {
                    const int c = rb - l;
                    if (c < 0 || c >= g.Lin) continue;
                    const float2 pv = P[(fl * g.Lin + c) * pstride + l * g.Nr + i];
                    sx += pv.x;
                    sy += pv.y;
                }
                const float2 zn = make_float2(yv[u].x - sx + bb[u] * zo[u].x, yv[u].y - sy + bb[u] * zo[u].y);     // scamp.py:48
                w.Z[f * g.n + o] = zn;
                w.Zs[f * (long long)a.zs_stride + o] = cdiv_real(zn, ph[u]);                                     // z / phi_use
            }
        }
    }
    if (MODE == 1 && MSEC > 0 && tid < a.FR && row_f[tid * g.Lin] >= 0) {
        // the tile holds every column block of its frames: retire them here (scamp.py:105); `notclose` was raised by this CTA's rows
        const long long f = f0 + tid;
        w.iters[f] = a.t + 1;
        if (g.early_exit && !w.notclose[f]) w.active[f] = 0;
        w.notclose[f] = 0;
    }
    __syncthreads();
    if (warp == kEpiWarps + kCvtWarps + 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
    }
}

// ---- design operand, split once per call ------------------------------------------------------------------------------------
// taps: [Lh][Nr][Nt] complex64.  Output layout per (chunk, stage): [4 planes: re_hi, re_lo, im_hi, im_lo][4 K chunks][brows][4 floats].
//   mode 0: row o = l Nr + i, reduction index k = j (column of the tap);              value = T_l[i][j]
//   mode 1: row = output column j of the chunk, reduction index k = l Kp + i (Kp = Nr padded to a multiple of KS);  value = conj(T_l[i][j])
__global__ void scamp_st_split_kernel(const float2* __restrict__ taps, int Lh, int Nr, int Nt, int mode, int brows, int chunks, int nks, int kp,
                                      float* __restrict__ out) {
    const long long total = (long long)chunks * nks * 4 * brows * 4;       // (chunk, stage, K chunk, row, element)
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int el = (int)(e % 4);
        const int row = (int)((e / 4) % brows);
        const int kc = (int)((e / (4LL * brows)) % 4);
        const int ks = (int)((e / (16LL * brows)) % nks);
        const int ch = (int)(e / (16LL * brows * nks));
        const int k = ks * KS + kc * 4 + el;
        float2 v = make_float2(0.f, 0.f);
        if (mode == 0) {
            const int l = row / Nr, i = row % Nr;
            if (row < Lh * Nr && k < Nt) v = taps[((size_t)l * Nr + i) * Nt + k];
        } else {
            const int j = ch * brows + row;
            const int l = k / kp, i = k % kp;
            if (l < Lh && i < Nr && j < Nt) {
                v = taps[((size_t)l * Nr + i) * Nt + j];
                v.y = -v.y;
            }
        }
        const float rh = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u), ih = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
        const size_t stage = ((size_t)ch * nks + ks) * (size_t)(16 * brows * 4);        // floats per stage
        const size_t at = stage + ((size_t)kc * brows + row) * 4 + el;
        const size_t plane = (size_t)4 * brows * 4;                                      // floats per plane
        out[at] = rh;
        out[at + plane] = v.x - rh;
        out[at + 2 * plane] = ih;
        out[at + 3 * plane] = v.y - ih;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

}  // namespace

static int round_up(int v, int m) { return (v + m - 1) / m * m; }

ScampStPlan scamp_st_plan(const Geom& g, int Lh) {
    ScampStPlan p{};
    p.ok = false;
    if (Lh < 1 || g.Lin < 1 || g.Lin > TM || (g.Nr & 1) || g.Lout < g.Lin || g.Lout > g.Lin + Lh - 1) return p;
    if (g.n != g.Lout * g.Nr || g.N != g.Lin * g.Nt) return p;
    p.brows0 = round_up(Lh * g.Nr, 32);
    if (p.brows0 > 128) return p;                              // P staging and TMEM: N of the residual MMAs <= 128
    p.brows1 = g.Nt >= 64 ? 64 : round_up(g.Nt, 32);        // one 64-column pair per chunk: 16 KiB design stages, two accumulator buffers of 128 columns
    p.chunks1 = (g.Nt + p.brows1 - 1) / p.brows1;
    p.nks0 = (g.Nt + KS - 1) / KS;
    p.kpb = (g.Nr + KS - 1) / KS;
    p.nks1 = Lh * p.kpb;
    p.FR = TM / g.Lin;
    p.zs_stride = (g.Lin + Lh - 1) * g.Nr;
    p.bplane_bytes0 = (size_t)p.nks0 * 16 * p.brows0 * 16;
    p.bplane_bytes1 = (size_t)p.chunks1 * p.nks1 * 16 * p.brows1 * 16;
    if (st_smem(p.brows0, 0).total > kSmemMax || st_smem(p.brows1, 1).total > kSmemMax) return p;
    p.ok = encode_tiled() != nullptr;
    return p;
}

int scamp_st_prepare(const Geom& g, const ScampStPlan& p, int Lh, const float2* taps, unsigned char* bplanes0, unsigned char* bplanes1,
                     cudaStream_t stream) {
    scamp_st_split_kernel<<<256, 256, 0, stream>>>(taps, Lh, g.Nr, g.Nt, 0, p.brows0, 1, p.nks0, 0, (float*)bplanes0);
    scamp_st_split_kernel<<<256, 256, 0, stream>>>(taps, Lh, g.Nr, g.Nt, 1, p.brows1, p.chunks1, p.nks1, p.kpb * KS, (float*)bplanes1);
    count_launch();
    count_launch();
    return check_cuda(cudaGetLastError(), "scamp_st_split_kernel launch");
}

// every symbol in {0, +-1, +-j}: one exact float32 product per exponent
static bool is_exact_alphabet(const DevAlphabet& al) {
    for (int k = 0; k < al.K; ++k) {
        const double ar = al.re[k] < 0 ? -al.re[k] : al.re[k], ai = al.im[k] < 0 ? -al.im[k] : al.im[k];
        if (!((ar == 0.0 || ar == 1.0) && (ai == 0.0 || ai == 1.0) && !(ar == 1.0 && ai == 1.0))) return false;
    }
    return true;
}

// exactly the four symbols 1, j, -1, -j in any order (the reference's QPSK table, config.py:93)
static bool is_axis4_alphabet(const DevAlphabet& al) {
    if (al.K != 4) return false;
    int seen = 0;
    for (int k = 0; k < 4; ++k) {
        const double r = al.re[k], i = al.im[k];
        if (r == 1.0 && i == 0.0) seen |= 1;
        else if (r == -1.0 && i == 0.0) seen |= 2;
        else if (r == 0.0 && i == 1.0) seen |= 4;
        else if (r == 0.0 && i == -1.0) seen |= 8;
        else return false;
    }
    return seen == 15;
}

bool scamp_st_can_fuse(const Geom& g, const DevAlphabet& al) {
    return (g.M == 8 || g.M == 16 || g.M == 32 || g.M == 64) && g.Nt == g.Na * g.M && al.K >= 1 && al.K <= AMPSM_MAX_K;
}

int scamp_st_gemm(int mode, const ScampWs& w, const Geom& g, const ScampStPlan& p, int Lh, const unsigned char* bplanes, const float2* y,
                  long long F, const DevAlphabet& al, bool fused, const float* W, float sigma2, const float* sigma2_pf, int t, cudaStream_t stream) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available"); return AMPSM_ENOFIT; }
    CUtensorMap map;
    CUresult r;
    if (mode == 0) {
        const cuuint64_t dims[2] = {(cuuint64_t)2 * g.Nt, (cuuint64_t)F * g.Lin};
        const cuuint64_t strides[1] = {(cuuint64_t)g.Nt * 8};
        const cuuint32_t box[2] = {(cuuint32_t)KS * 2, (cuuint32_t)(p.FR * g.Lin)};
        const cuuint32_t es[2] = {1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)w.Xh, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        // zs[f] as (element in row block, row block, frame); the tile of tap l is the box at block coordinate l: rows (f, c) read
        // block c + l.  Blocks >= Lout (truncated channels) are zero padding of the workspace; elements >= Nr are out of bounds = 0.
        const cuuint64_t dims[3] = {(cuuint64_t)2 * g.Nr, (cuuint64_t)(g.Lin + Lh - 1), (cuuint64_t)F};
        const cuuint64_t strides[2] = {(cuuint64_t)g.Nr * 8, (cuuint64_t)p.zs_stride * 8};
        const cuuint32_t box[3] = {(cuuint32_t)KS * 2, (cuuint32_t)g.Lin, (cuuint32_t)p.FR};
        const cuuint32_t es[3] = {1, 1, 1};
        r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)w.Zs, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) for SCAMP mode %d", (int)r, mode); return AMPSM_ENOFIT; }
    StArgs a{};
    a.W = W; a.sigma2 = sigma2; a.sigma2_pf = sigma2_pf; a.t = t;
    a.w = w; a.g = g; a.al = al; a.y = y; a.bplanes = bplanes; a.F = F; a.Lh = Lh; a.FR = p.FR; a.zs_stride = p.zs_stride; a.kpb = p.kpb;
    a.brows = mode == 0 ? p.brows0 : p.brows1;
    if (const char* d = getenv("AMPSM_ST_DEBUG")) a.dbg = atoi(d);
    a.chunks = mode == 0 ? 1 : p.chunks1;
    a.nks = mode == 0 ? p.nks0 : p.nks1;
    if (const char* r = getenv(mode == 0 ? "AMPSM_ST_RINGS0" : "AMPSM_ST_RINGS1")) sscanf(r, "%d,%d", &a.ring_raw, &a.ring_b);
    const int smem = st_smem(a.brows, mode, a.ring_raw, a.ring_b).total;
    const unsigned grid = (unsigned)((F + p.FR - 1) / p.FR);
    auto run = [&](auto kern, const char* what) -> int {
        if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem), what)) return e;
        kern<<<grid, kThreads, smem, stream>>>(a, map);
        return 0;
    };
    int e = 0;
    const bool exact = fused && is_exact_alphabet(al) && !getenv("AMPSM_SCAMP_NO_EXACT");
    const bool axis4 = exact && is_axis4_alphabet(al) && !getenv("AMPSM_SCAMP_NO_AXIS4");
    if (mode == 0) e = run(scamp_st_kernel<0, 0, 0>, "cudaFuncSetAttribute(scamp_st<0>)");
    else if (!fused) e = run(scamp_st_kernel<1, 0, 0>, "cudaFuncSetAttribute(scamp_st<1>)");
#define AMPSM_ST_CASE(MM) \
    else if (g.M == MM) e = axis4 ? run(scamp_st_kernel<1, MM, 2>, "cudaFuncSetAttribute(scamp_st<1, " #MM ", axis4>)") \
                          : exact ? run(scamp_st_kernel<1, MM, 1>, "cudaFuncSetAttribute(scamp_st<1, " #MM ", exact>)") \
                                  : run(scamp_st_kernel<1, MM, 0>, "cudaFuncSetAttribute(scamp_st<1, " #MM ">)");
    AMPSM_ST_CASE(64)
    AMPSM_ST_CASE(32)
    AMPSM_ST_CASE(16)
    AMPSM_ST_CASE(8)
#undef AMPSM_ST_CASE
    if (e) return e;
    count_launch();
    return check_cuda(cudaGetLastError(), "scamp_st_kernel launch");
}

}  // namespace ampsm
