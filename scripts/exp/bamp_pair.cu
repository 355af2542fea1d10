// KEPT AS EVIDENCE, NOT BUILT: measured 15 % slower than the one-warp kernel on B200; it was removed from the
// product library in round 2.  To try it again: copy it to amp-sparc-spatialmodulation_b200/csrc/ and restore its dispatch in cabi.cu.
// Register-resident BAMP kernel, TWO WARPS PER FRAME (one 64-thread CTA per frame, 64 x 32 shapes).
//
// bamp_fast.cu keeps a whole 32 x 64 channel matrix in the registers of ONE warp (192 registers of tile per lane),
// which caps an SM at 8 warps: ncu showed that kernel latency-bound on the serial part of an iteration (48 % issue
// slots, 47 % FMA pipe).  Here a frame is split by COLUMNS over two warps: warp w owns columns [32w, 32w+32) of H and
// of |H|^2 (96 registers of tile per lane), so an SM holds 6 frames = 12 warps and the per-warp serial chain is half
// as long.  Per iteration (bamp.py:59-64):
//   1. row pass     each warp: partial  v = |H|^2 var,  H xhat  over its own 32 columns        -> xrow (shared)
//      barrier A    (also carries the allclose vote of the previous iteration, bamp.py:140)
//   2. row update   BOTH warps reduce the 8 partials of row `lane` and update z, u (bamp.py:60-61) redundantly and
//                   bit-identically: nothing has to travel back
//   3. column pass  each warp: |H|^2^T (1/u), H^H((y-z)/u) for its own columns, reduced inside the warp
//   4. denoiser     one column per lane (bamp.py:66-77); a section that spans both warps (M = 64) is stitched with
//                   ONE exchange of {warp maximum, warp sum}: each warp's exponentials are relative to its own
//                   maximum and rescaled by 2^(own max - section max)
//      barrier B    (the exchange; also separates this iteration's xrow reads from the next row pass)
// The lanes of a warp form an 8 x 4 grid (8 row groups x 4 column groups, 4 x 8 tile each): 8 partials per row and
// 8 per column, all crossing lanes through float4 arrays in shared memory whose rows are padded to 9 entries: conflict-
// free both ways and every address in the loop is a per-lane base plus an immediate.
// All mat-vec FMAs are packed FFMA2 on the natural (re,im) pairs as in bamp_fast.cu.  The next frame's H and y are
// prefetched by a 1-D bulk TMA copy into the CTA's staging buffer as soon as both warps hold their tiles.
// Loss (MAP decision, counters: loss.py:67-179, 282-302) is fused as in bamp_fast.cu.
#include <cstdlib>

#include "fastops.cuh"

namespace ampsm {

namespace {

constexpr int kPairThreads = 64;
constexpr int kPairCtasPerSm = 6;

struct PairSmem {   // byte offsets in the CTA's dynamic shared memory (n = 32, N = 64)
    static constexpr int stage = 0;                      // H [32][64] complex64 + y [32]
    static constexpr int stage_bytes = 32 * 64 * 8 + 32 * 8;
    static constexpr int xrow = stage + stage_bytes;     // float4 [32 rows][8 partials + 1 pad]          (both warps)
    static constexpr int xcol = xrow + 32 * 9 * 16;      // float4 [2][32 columns][8 partials + 1 pad]    (per warp)
    static constexpr int colvec = xcol + 2 * 32 * 9 * 16;   // float4 [2][32] {xx,xx,xy,xy}
    static constexpr int varvec = colvec + 2 * 32 * 16;  // float  [2][32]
    static constexpr int rowvec = varvec + 2 * 32 * 4;   // float4 [2][32] {gx,gy,gy,-gx}
    static constexpr int wvec = rowvec + 2 * 32 * 16;    // float2 [2][32] {1/u,1/u}
    static constexpr int misc = wvec + 2 * 32 * 8;       // float2 zx[2]; 16 B pick[2]; float4 trj[2]
    static constexpr int cnt = misc + 128;               // u64 [2][16]
    static constexpr int mbar = cnt + 2 * 16 * 8;
    static constexpr int total = mbar + 16;
};
static_assert((PairSmem::total + 1024) * kPairCtasPerSm <= 233472, "the frames of one SM must fit its shared memory");

template <int W>
__device__ __forceinline__ float seg_max(float m) {
    if constexpr (W == 32) {
        float r;
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(m));
        return r;
    } else {
#pragma unroll
        for (int o = W / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        return m;
    }
}
// Z = sum over the W-lane segment; others = Z - own WITHOUT cancellation (what a lane receives in the butterfly)
template <int W>
__device__ __forceinline__ void seg_sum_excl(float own, float& Z, float& others) {
    float part = own, recv = 0.f;
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) {
        const float r = __shfl_xor_sync(0xffffffffu, part, o);
        recv += r;
        part += r;
    }
    Z = part;
    others = recv;
}

template <int M_, int K_, bool GRID>
__global__ void __launch_bounds__(kPairThreads, kPairCtasPerSm) bamp_pair_kernel(const __grid_constant__ BampArgs a) {
    constexpr int n = 32, N = 64, L_ = N / M_;
    constexpr int W = M_ >= 32 ? 32 : M_;                 // lanes of one section inside a warp
    constexpr bool SPLIT = M_ == 64;                      // the section spans both warps
    static_assert(M_ == 64 || M_ == 32 || M_ == 16 || M_ == 8, "section sizes of the 64-column shapes");
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int la = lane >> 2, lb = lane & 3;
    const float2* stH = reinterpret_cast<const float2*>(smem + PairSmem::stage);
    const float2* stY = reinterpret_cast<const float2*>(smem + PairSmem::stage + n * N * 8);
    float4* xrow = reinterpret_cast<float4*>(smem + PairSmem::xrow);
    float4* xcol = reinterpret_cast<float4*>(smem + PairSmem::xcol) + w * 288;
    float4* colvec = reinterpret_cast<float4*>(smem + PairSmem::colvec) + w * 32;
    float* varvec = reinterpret_cast<float*>(smem + PairSmem::varvec) + w * 32;
    float4* rowvec = reinterpret_cast<float4*>(smem + PairSmem::rowvec) + w * 32;
    float2* wvec = reinterpret_cast<float2*>(smem + PairSmem::wvec) + w * 32;
    float2* zx = reinterpret_cast<float2*>(smem + PairSmem::misc);
    Pick* pickx = reinterpret_cast<Pick*>(smem + PairSmem::misc + 16);
    float4* trj = reinterpret_cast<float4*>(smem + PairSmem::misc + 48);
    unsigned long long* cnt = reinterpret_cast<unsigned long long*>(smem + PairSmem::cnt) + w * 16;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + PairSmem::mbar);

    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    constexpr uint32_t kHBytes = n * N * 8, kYBytes = n * 8;
    const int colg = w * 32 + lane;                       // the column this lane owns

    if (threadIdx.x == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    if (lane < 16) cnt[lane] = 0ull;
    __syncthreads();
    auto prefetch = [&](long long f) {
        mbar_expect_tx(mbar, kHBytes + kYBytes);
        tma_load_1d(smem + PairSmem::stage, a.H + f * a.H_stride, kHBytes, mbar);
        tma_load_1d(smem + PairSmem::stage + kHBytes, a.y + f * n, kYBytes, mbar);
    };
    long long f = blockIdx.x;
    if (f < a.frames && threadIdx.x == 0) prefetch(f);
    uint32_t phase = 0;

    for (; f < a.frames; f += gridDim.x) {
        mbar_wait(mbar, phase);
        phase ^= 1u;
        // ---- staging buffer -> registers: lane (la,lb) of warp w takes rows la*4+i, columns 32w + (t*4+lb)*2 + e
        pair_t Hp[4][8], Pp[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2* row = stH + (la * 4 + i) * N + w * 32;
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(row + (t * 4 + lb) * 2);
                Hp[i][2 * t] = v.x;
                Hp[i][2 * t + 1] = v.y;
                float a0, a1, b0, b1;
                unpack2(fmul2(v.x, v.x), a0, a1);
                unpack2(fmul2(v.y, v.y), b0, b1);
                Pp[i][t] = pack2(a0 + a1, b0 + b1);          // |H|^2 (bamp.py:18) of the two adjacent columns
            }
        }
        const float2 yv = stY[lane];
        const float sigma2 = a.sigma2_pf ? a.sigma2_pf[f] : a.sigma2;
        if (a.io.x_true) {   // the Loss epilogue reads these once, right after the last iteration: start the fetch now
            if (threadIdx.x < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.io.x_true + f * N + threadIdx.x * 16));
            if (threadIdx.x == 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.io.idx_true + f * L_));
            if (threadIdx.x == 33) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.io.sym_true + f * L_));
        }
        // state (bamp.py:20-25): lane r of BOTH warps keeps z_r, u_r; lane c of warp w keeps xhat, var, xmap of its column
        float zr = yv.x, zi = yv.y, u = sigma2;
        float xhr = 0.f, xhi = 0.f, var_own = 1.0f, xmr = 0.f, xmi = 0.f;
        colvec[lane] = make_float4(0.f, 0.f, 0.f, 0.f);
        varvec[lane] = 1.0f;
        __syncthreads();     // both warps hold their tiles: the staging buffer is free for the next frame
        if (threadIdx.x == 0) {
            const long long nf = f + gridDim.x;
            if (nf < a.frames) prefetch(nf);
        }

        int t_done = 0;
        bool close = false;
        for (int it = 0; it < g.max_iters; ++it) {
            // ================= row pass over the warp's own columns: v = |H|^2 var, H xhat (bamp.py:59-60) =========
            {
                pair_t A[4], B[4], V[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const int c0 = (t * 4 + lb) * 2;
                    const ulonglong2 x0 = *reinterpret_cast<const ulonglong2*>(&colvec[c0]);        // {xx,xx | xy,xy}
                    const ulonglong2 x1 = *reinterpret_cast<const ulonglong2*>(&colvec[c0 + 1]);
                    const pair_t vp = *reinterpret_cast<const pair_t*>(&varvec[c0]);                // {var_c, var_c+1}
                    if (t == 0) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) A[i] = fmul2(Hp[i][0], x0.x);
#pragma unroll
                        for (int i = 0; i < 4; ++i) B[i] = fmul2(Hp[i][0], x0.y);
#pragma unroll
                        for (int i = 0; i < 4; ++i) V[i] = fmul2(Pp[i][0], vp);
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; ++i) A[i] = ffma2(Hp[i][2 * t], x0.x, A[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) B[i] = ffma2(Hp[i][2 * t], x0.y, B[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) V[i] = ffma2(Pp[i][t], vp, V[i]);
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i) A[i] = ffma2(Hp[i][2 * t + 1], x1.x, A[i]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) B[i] = ffma2(Hp[i][2 * t + 1], x1.y, B[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = la * 4 + i;
                    float al_, ah_, bl_, bh_, vl_, vh_;
                    unpack2(A[i], al_, ah_);
                    unpack2(B[i], bl_, bh_);
                    unpack2(V[i], vl_, vh_);
                    xrow[row * 9 + w * 4 + lb] = make_float4(vl_ + vh_, al_ - bh_, bl_ + ah_, vl_);
                }
            }
            // barrier A: the row partials of both warps are in place; the vote is last iteration's allclose (bamp.py:140)
            const int all_close = __syncthreads_and(close ? 1 : 0);
            if (g.early_exit && all_close) break;
            // ================= row update, redundantly in both warps (bamp.py:60-61) =================
            {
                float4 p[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) p[j] = xrow[lane * 9 + j];
                const float sv = ((p[0].x + p[1].x) + (p[2].x + p[3].x)) + ((p[4].x + p[5].x) + (p[6].x + p[7].x));
                const float sr = ((p[0].y + p[1].y) + (p[2].y + p[3].y)) + ((p[4].y + p[5].y) + (p[6].y + p[7].y));
                const float si = ((p[0].z + p[1].z) + (p[2].z + p[3].z)) + ((p[4].z + p[5].z) + (p[6].z + p[7].z));
                // z = Hx - v (y - z)/u_old ; u = v + sigma2 ; operands of the column pass (bamp.py:60-63)
                const float ru = fast_rcp(u);
                const float znr = sr - sv * (yv.x - zr) * ru, zni = si - sv * (yv.y - zi) * ru;
                u = sv + sigma2;
                const float rn = fast_rcp(u);
                zr = znr;
                zi = zni;
                const float gx = (yv.x - zr) * rn, gy = (yv.y - zi) * rn;
                rowvec[lane] = make_float4(gx, gy, gy, -gx);        // operand pairs (gx,gy), (gy,-gx)
                wvec[lane] = make_float2(rn, rn);
            }
            __syncwarp();
            // ================= column pass: cov = 1/(|H|^2^T 1/u), H^H((y-z)/u) (bamp.py:62-63) =================
            float cov;
            {
                // columns in two halves: 20 accumulator registers instead of 40 next to the 96-register tile
#pragma unroll
                for (int c0 = 0; c0 < 8; c0 += 4) {
                    pair_t A[4], B[4], C[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = la * 4 + i;
                        const ulonglong2 gq = *reinterpret_cast<const ulonglong2*>(&rowvec[row]);   // {gx,gy | gy,-gx}
                        const pair_t wp = *reinterpret_cast<const pair_t*>(&wvec[row]);             // {1/u, 1/u}
                        if (i == 0) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) A[c] = fmul2(Hp[0][c0 + c], gq.x);
#pragma unroll
                            for (int c = 0; c < 4; ++c) B[c] = fmul2(Hp[0][c0 + c], gq.y);
#pragma unroll
                            for (int c = 0; c < 2; ++c) C[c] = fmul2(Pp[0][c0 / 2 + c], wp);
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) A[c] = ffma2(Hp[i][c0 + c], gq.x, A[c]);
#pragma unroll
                            for (int c = 0; c < 4; ++c) B[c] = ffma2(Hp[i][c0 + c], gq.y, B[c]);
#pragma unroll
                            for (int c = 0; c < 2; ++c) C[c] = ffma2(Pp[i][c0 / 2 + c], wp, C[c]);
                        }
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int col = (((c0 + c) >> 1) * 4 + lb) * 2 + (c & 1);      // local column of tile column c0+c
                        float lo, hi, cl, ch;
                        unpack2(A[c], lo, hi);
                        const float cr = lo + hi;
                        unpack2(B[c], lo, hi);
                        const float ci = lo + hi;
                        unpack2(C[c >> 1], cl, ch);
                        xcol[col * 9 + la] = make_float4((c & 1) ? ch : cl, cr, ci, cr);
                    }
                }
                __syncwarp();
                float4 p[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) p[j] = xcol[lane * 9 + j];
                const float sc = ((p[0].x + p[1].x) + (p[2].x + p[3].x)) + ((p[4].x + p[5].x) + (p[6].x + p[7].x));
                const float sr = ((p[0].y + p[1].y) + (p[2].y + p[3].y)) + ((p[4].y + p[5].y) + (p[6].y + p[7].y));
                const float si = ((p[0].z + p[1].z) + (p[2].z + p[3].z)) + ((p[4].z + p[5].z) + (p[6].z + p[7].z));
                cov = fast_rcp(sc);
                xmr = fmaf(cov, sr, xhr);
                xmi = fmaf(cov, si, xhi);
            }
            // ================= denoiser (bamp.py:66-77), tau = cov/2, one column per lane =================
            float xr, xi, vn;
            {
                const float rt = fast_rcp(cov * 0.5f);
                const float q_r = xmr * rt, q_i = xmi * rt;
                float S0, S1r, S1i, Zl, others_l, lmax_w;
                // GRID (the reference's 16-QAM table, see bamp_fast.cu): e_k = Er[a_k] Ei[b_k] on the 4 x 4 level grid
                float Er[GRID ? 4 : 1], Ei[GRID ? 4 : 1], a0 = 0.f, b0 = 0.f, e13 = 0.f, e20 = 0.f;
                float ek[GRID ? 1 : K_];
                if constexpr (GRID) {
                    const DevGrid& G = a.grid;
                    const bool rp = q_r >= 0.f, ip = q_i >= 0.f;
                    const double lmd = (double)q_r * (rp ? G.lr2[3] : G.lr2[0]) + (double)q_i * (ip ? G.li2[3] : G.li2[0]);
                    lmax_w = seg_max<W>((float)lmd);
                    const float off = (float)(lmd - (double)lmax_w);       // <= 0 up to rounding
                    float a1 = 0.f, b1 = 0.f;
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        Er[l] = fast_ex2(q_r * (rp ? G.dpos_r[l] : G.dneg_r[l]));
                        Ei[l] = fast_ex2(fmaf(q_i, ip ? G.dpos_i[l] : G.dneg_i[l], off));
                        a0 += Er[l];
                        a1 = fmaf(G.lrf[l], Er[l], a1);
                        b0 += Ei[l];
                        b1 = fmaf(G.lif[l], Ei[l], b1);
                    }
                    e13 = Er[1] * Ei[3];
                    e20 = Er[2] * Ei[0];
                    S0 = fmaf(a0, b0, e13 - e20);
                    S1r = fmaf(a1, b0, fmaf(G.lrf[1], e13, -G.lrf[2] * e20));
                    S1i = fmaf(a0, b1, fmaf(G.lif[3], e13, -G.lif[0] * e20));
                } else {
                    const double qr = (double)q_r, qi = (double)q_i;
                    float m = -INFINITY;
#pragma unroll
                    for (int k = 0; k < K_; ++k) m = fmaxf(m, fmaf(q_r, al.ref[k], q_i * al.imf[k]));
                    lmax_w = seg_max<W>(m);                                // a common shift, nothing else
                    const double shift = (double)lmax_w;
                    S0 = S1r = S1i = 0.f;
#pragma unroll
                    for (int k = 0; k < K_; ++k) {
                        const double x = fma(qr, al.re[k], qi * al.im[k]);
                        const float e = fast_ex2((float)(x - shift) * 1.4426950408889634f);
                        ek[k] = e;
                        S0 += e;
                        S1r = fmaf(al.ref[k], e, S1r);
                        S1i = fmaf(al.imf[k], e, S1i);
                    }
                }
                seg_sum_excl<W>(S0, Zl, others_l);
                // barrier B: stitch a section that spans both warps; for in-warp sections it only orders the xrow reuse
                float rz, rzo, others;
                if constexpr (SPLIT) {
                    if (lane == 0) zx[w] = make_float2(lmax_w, Zl);
                    __syncthreads();
                    const float2 o = zx[w ^ 1];
                    const float m = fmaxf(lmax_w, o.x);
                    constexpr float unit = GRID ? 1.0f : 1.4426950408889634f;   // GRID exponents are already in log2 units
                    const float so = fast_ex2((lmax_w - m) * unit), st = fast_ex2((o.x - m) * unit);
                    const float Zo = o.y * st;
                    const float Z = fmaf(Zl, so, Zo);
                    others = fmaf(others_l, so, Zo);
                    rz = fast_rcp(Z);
                    rzo = so * rz;
                } else {
                    __syncthreads();
                    rz = rzo = fast_rcp(Zl);
                    others = others_l;
                }
                xr = S1r * rzo;
                xi = S1i * rzo;
                // two-term variance (bamp.py:74-76)
                float spread;
                if constexpr (GRID) {
                    const DevGrid& G = a.grid;
                    float dr = 0.f, di = 0.f, er2[4], ei2[4];
#pragma unroll
                    for (int l = 0; l < 4; ++l) {
                        const float er = xr - G.lrf[l], ei = xi - G.lif[l];
                        er2[l] = er * er;
                        ei2[l] = ei * ei;
                        dr = fmaf(er2[l], Er[l], dr);
                        di = fmaf(ei2[l], Ei[l], di);
                    }
                    spread = fmaf(dr, b0, a0 * di);
                    spread = fmaf(er2[1] + ei2[3], e13, spread);
                    spread = fmaf(-(er2[2] + ei2[0]), e20, spread);
                } else {
                    spread = 0.f;
#pragma unroll
                    for (int k = 0; k < K_; ++k) {
                        const float dr = xr - al.ref[k], di = xi - al.imf[k];
                        spread = fmaf(fmaf(dr, dr, di * di), ek[k], spread);
                    }
                }
                vn = fmaf(fmaf(xr, xr, xi * xi), others * rz, spread * rzo);
            }
            // exit test on var (bamp.py:140; voted at the next barrier A), publish the estimate for the next row pass
            close = __all_sync(0xffffffffu, fabsf(vn - var_own) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, var_own))));
            xhr = xr;
            xhi = xi;
            var_own = vn;
            colvec[lane] = make_float4(xr, xr, xi, xi);
            varvec[lane] = vn;
            if (a.traj) {
                float s_mse = 0.f;
                if (a.io.x_true) {
                    const float2 xt = a.io.x_true[f * N + colg];
                    s_mse = (xr - xt.x) * (xr - xt.x) + (xi - xt.y) * (xi - xt.y);
                }
                const float s_tau = warp_sum(cov), s_var = warp_sum(vn);
                s_mse = warp_sum(s_mse);
                if (lane == 0) trj[w] = make_float4(s_tau, s_var, s_mse, 0.f);
                __syncthreads();
                if (threadIdx.x == 0) {
                    float* tr = a.traj + (f * g.max_iters + it) * 3;
                    tr[0] = (trj[0].x + trj[1].x) / N;
                    tr[1] = (trj[0].y + trj[1].y) / N;
                    tr[2] = (trj[0].z + trj[1].z) / N;
                }
            }
            __syncwarp();
            t_done = it + 1;
        }

        // ================= outputs =================
        if (a.xmap) a.xmap[f * N + colg] = make_float2(xmr, xmi);
        if (a.xmmse) a.xmmse[f * N + colg] = make_float2(xhr, xhi);
        if (a.var) a.var[f * N + colg] = var_own;
        if (threadIdx.x == 0) {
            if (a.traj)
                for (int it = t_done; it < g.max_iters; ++it)
                    for (int q = 0; q < 3; ++q)
                        a.traj[(f * g.max_iters + it) * 3 + q] = a.traj[(f * g.max_iters + t_done - 1) * 3 + q];
            if (a.iters) a.iters[f] = t_done;
            cnt[C_FRAMES] += 1;
            cnt[C_ITERS] += t_done;
        }

        // ================= Loss: MAP decision + counters (loss.py:282-302, 67-179), Lin = 1 =================
        if (a.io.x_true) {
            const int m = colg % M_, sec = colg / M_;
            // in-order scan over k (flat index increases with k): the first maximum wins; a NaN wins once and sticks
            const double xr = (double)xmr, xi = (double)xmi;
            double bv = __dadd_rn(__dmul_rn(xr, al.re[0]), __dmul_rn(xi, al.im[0]));
            int bk = 0;
#pragma unroll
            for (int k = 1; k < K_; ++k) {
                const double v = __dadd_rn(__dmul_rn(xr, al.re[k]), __dmul_rn(xi, al.im[k]));
                const bool upd = (bv == bv) & ((v > bv) | (v != v));
                bv = upd ? v : bv;
                bk = upd ? k : bk;
            }
            Pick b{bv, m * K_ + bk};
            const bool nan_seen = (xmr != xmr) || (xmi != xmi);
#pragma unroll
            for (int o = W / 2; o > 0; o >>= 1) {
                Pick other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.idx, o)};
                if (pick_better(other, b)) b = other;
            }
            if constexpr (SPLIT) {
                if (lane == 0) pickx[w] = b;
                __syncthreads();
                const Pick other = pickx[w ^ 1];
                if (pick_better(other, b)) b = other;
            }
            const int dec_ant = b.idx / K_, k = b.idx % K_;
            const float2 xt = a.io.x_true[f * N + colg];
            const float2 h = (m == dec_ant) ? make_float2((float)al.re[k], (float)al.im[k]) : make_float2(0.f, 0.f);
            const bool wrong = (h.x != xt.x) || (h.y != xt.y);
            const float dr = xhr - xt.x, di = xhi - xt.y;
            double sq = (double)dr * dr + (double)di * di;
            unsigned long long packed = 0ull;
            if (m == 0) {   // one lane per section books the label counters
                const long long ih = (g.frame_base + f) * (long long)N + sec * M_ + dec_ant;
                const long long itrue = a.io.idx_true[f * L_ + sec];
                const long long sh = al.gray[k], st = a.io.sym_true[f * L_ + sec];
                const unsigned long long imask = g.index_bits_kept >= 64 ? ~0ull : ((1ull << g.index_bits_kept) - 1ull);
                packed = (unsigned long long)(ih != itrue) | ((unsigned long long)(sh != st) << 12) |
                         ((unsigned long long)__popcll((unsigned long long)(ih ^ itrue) & imask) << 24) |
                         ((unsigned long long)__popcll((unsigned long long)(sh ^ st) & ((1ull << al.sbits) - 1ull)) << 44);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
            sq = warp_sum(sq);
            const int any_wrong = __syncthreads_or(wrong ? 1 : 0);
            const int any_nan = __syncthreads_or(nan_seen ? 1 : 0);
            if (lane == 0) {
                cnt[C_INDEX_ERR] += packed & 0xfffull;
                cnt[C_SYMBOL_ERR] += (packed >> 12) & 0xfffull;
                cnt[C_INDEX_BIT] += (packed >> 24) & 0xfffffull;
                cnt[C_SYMBOL_BIT] += packed >> 44;
                reinterpret_cast<double*>(cnt)[12] += sq;
                if (w == 0) {
                    cnt[C_FRAME_ERR] += any_wrong ? 1 : 0;               // Lin = 1: one time slot per frame
                    cnt[C_NAN_FRAMES] += any_nan ? 1 : 0;
                }
            }
        }
        __syncthreads();     // the frame is done: shared state may be re-initialised
    }

    // ---- flush the warps' counters
    __syncthreads();
    if (a.io.counters && lane == 0) {
        unsigned long long* out = a.io.counters;
        const int plain[] = {C_FRAMES, C_INDEX_ERR, C_SYMBOL_ERR, C_INDEX_BIT, C_SYMBOL_BIT, C_ITERS, C_NAN_FRAMES};
        for (int k : plain)
            if (cnt[k]) atomicAdd(out + k, cnt[k]);
        if (cnt[C_FRAME_ERR]) {          // Lin = 1: the frame is its only, first, middle and last time slot
            const int slots[] = {C_FRAME_ERR, C_SLOT_ERR, C_SLOT_FIRST, C_SLOT_MID, C_SLOT_LAST};
            for (int k : slots) atomicAdd(out + k, cnt[C_FRAME_ERR]);
        }
        const double sq = reinterpret_cast<double*>(cnt)[12];
        if (sq != 0.0)
            for (int k = 0; k < 4; ++k) atomicAdd(reinterpret_cast<double*>(out) + C_SQERR + k, sq);
    }
}

template <int M_, int K_, bool GRID>
int launch_pair_shape(const BampArgs& a, cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = bamp_pair_kernel<M_, K_, GRID>;
    const size_t smem = PairSmem::total;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                           "cudaFuncSetAttribute(bamp_pair)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kPairThreads, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > a.frames) grid = a.frames;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kPairThreads, smem, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "bamp_pair_kernel launch");
}

}  // namespace

int launch_bamp_pair(const BampArgs& a, cudaStream_t stream) {
    const Geom& g = a.g;
    // the fused Loss epilogue assumes one time slot per frame; the staging path needs 16-byte aligned frames
    if (g.Lin != 1 || g.decision != 0 || g.shift_mode != 0 || g.n != 32 || g.N != 64) return AMPSM_ENOFIT;
    if ((reinterpret_cast<uintptr_t>(a.H) % 16) || (reinterpret_cast<uintptr_t>(a.y) % 16) ||
        (a.H_stride != 0 && ((size_t)a.H_stride * 8) % 16))
        return AMPSM_ENOFIT;
    const int K = a.al.K;
    BampArgs b = a;
    b.grid = make_grid(a.al);
    if (g.M == 64 && K == 16 && b.grid.ok && !getenv("AMPSM_NO_GRID")) return launch_pair_shape<64, 16, true>(b, stream);
#define AMPSM_PAIR(MM, KK) \
    if (g.M == MM && K == KK) return launch_pair_shape<MM, KK, false>(b, stream);
    AMPSM_PAIR(64, 16)    // C2 with the table-driven denoiser
    AMPSM_PAIR(64, 4)     // 64 x 32, QPSK
    AMPSM_PAIR(16, 4)     // 64 x 32, QPSK, Na = 4
    AMPSM_PAIR(32, 4)     // 64 x 32, QPSK, Na = 2
#undef AMPSM_PAIR
    return AMPSM_ENOFIT;
}

}  // namespace ampsm
