// Device helpers shared by the register-resident kernels (bamp_fast.cu, bamp_pair.cu, vamp_pair.cu): packed
// fp32x2 arithmetic, approximate MUFU wrappers, the product-grid view of an alphabet and the section reductions
// over the column-owner layout.
#pragma once
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

// Phase timing for development builds (-DAMPSM_CLK, scripts/build_clk.sh, scripts/phase_clocks.py): lane 0 of every warp
// accumulates the cycles between phase boundaries in `clkacc` (16 shared 32-bit slots per warp); the kernels add them to
// a global array that ampsm_debug_clocks() / ampsm_debug_clocks_vamp() return.  Compiled out of the shipped library.
#ifdef AMPSM_CLK
#define CLK_INIT() unsigned clk_last_ = clock()
#define CLK(p) do { const unsigned now_ = clock(); if (lane == 0) atomicAdd(&clkacc[p], now_ - clk_last_); clk_last_ = now_; } while (0)
#else
#define CLK_INIT() do {} while (0)
#define CLK(p) do {} while (0)
#endif

// torch.max / torch.min propagate NaN (vamp.py:76-77)
__device__ __forceinline__ float clampF(float v, float lo, float hi) {
    if (v != v) return v;
    return fminf(fmaxf(v, lo), hi);
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// reciprocal to <= 1 ulp without a slow path: MUFU.RCP + one Newton step (two FMAs).  Tried (round 2) for the VAMP scalars whose
// error the iteration amplifies (1 / (1 - alpha), 1 / (1 - dxdr), vamp.py:79-91) in place of fast_rcp: the number of frames that
// decide differently from the oracle did not move (270 / 96 / 62 / 32 of 10^4 at 5 / 10 / 15 / 20 dB against 278 / 94 / 73 / 32)
// and the kernel lost 1 %, so the VAMP kernels keep MUFU.RCP -- the flips come from summation order, not from the reciprocals.
__device__ __forceinline__ float rcp_ulp(float x) {
    const float r = fast_rcp(x);
    return fmaf(fmaf(-x, r, 1.0f), r, r);
}
// Packed fp32x2 arithmetic on 64-bit register pairs (FFMA2, sm_100).  The pairs are held in 64-bit containers so
// that ptxas keeps the H tile PACKED across the whole frame; with float2 values it re-assembles every operand pair
// with two MOVs per FFMA2 inside the iteration loop.
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t pack2(float lo, float hi) {
    pair_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(pair_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ pair_t ffma2(pair_t a, pair_t b, pair_t c) {
    pair_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ pair_t fmul2(pair_t a, pair_t b) {
    pair_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
// Scheduling tie: returns v unchanged (opaque_zero is a kernel parameter that is always 0, which ptxas cannot know), but as the
// result of an instruction that also reads the low half of `acc` -- so everything computed from the returned value is issued
// after the instruction that produced `acc`.  Used to spread a latency chain over a stream of independent FFMA2 (vamp_quad.cu).
__device__ __forceinline__ float chain_tie(float v, pair_t acc, unsigned opaque_zero) {
    unsigned r = __float_as_uint(v);
    const unsigned lo = (unsigned)acc;
    asm("lop3.b32 %0, %0, %1, %2, 0xF8;" : "+r"(r) : "r"(lo), "r"(opaque_zero));      // r | (lo & zero)
    return __uint_as_float(r);
}
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// Product-grid view of an alphabet: sym_k = (lr[a_k], li[b_k]) on an NGR x NGI grid, with `ncorr` grid points whose
// multiplicity in the table differs from one (the reference's 16-QAM list has one point twice and one missing,
// config.py:112).  exp(Re(q conj s)) then factorises into a real-part and an imaginary-part factor, so an antenna
// needs NGR + NGI exponentials instead of K.

inline DevGrid make_grid(const DevAlphabet& al) {
    DevGrid g{};
    double lr[AMPSM_MAX_K], li[AMPSM_MAX_K];
    int nr = 0, ni = 0;
    auto add = [](double* v, int& n, double x) {
        for (int i = 0; i < n; ++i)
            if (v[i] == x) return;
        v[n++] = x;
    };
    for (int k = 0; k < al.K; ++k) {
        add(lr, nr, al.re[k]);
        add(li, ni, al.im[k]);
    }
    if (nr > kGridMax || ni > kGridMax) return g;
    auto sort = [](double* v, int n) {
        for (int i = 0; i < n; ++i)
            for (int j = i + 1; j < n; ++j)
                if (v[j] < v[i]) { double t = v[i]; v[i] = v[j]; v[j] = t; }
    };
    sort(lr, nr);
    sort(li, ni);
    int cnt[kGridMax][kGridMax] = {};
    for (int k = 0; k < al.K; ++k) {
        int a = 0, b = 0;
        while (lr[a] != al.re[k]) ++a;
        while (li[b] != al.im[k]) ++b;
        cnt[a][b]++;
    }
    int nc = 0;
    for (int a = 0; a < nr; ++a)
        for (int b = 0; b < ni; ++b)
            if (cnt[a][b] != 1) {
                if (nc == kGridCorrMax) return g;
                g.ca[nc] = a; g.cb[nc] = b; g.cw[nc] = (float)(cnt[a][b] - 1);
                ++nc;
            }
    if (nr != kGridMax || ni != kGridMax) return g;          // only full 4 x 4 grids take the separable path for now
    g.nr = nr; g.ni = ni; g.ncorr = nc;
    const double log2e = 1.4426950408889634074;
    for (int i = 0; i < kGridMax; ++i) {
        g.lr2[i] = lr[i] * log2e; g.li2[i] = li[i] * log2e;
        g.lr2f[i] = (float)g.lr2[i]; g.li2f[i] = (float)g.li2[i];
        g.lr2l[i] = (float)(g.lr2[i] - (double)g.lr2f[i]); g.li2l[i] = (float)(g.li2[i] - (double)g.li2f[i]);
        g.lrf[i] = (float)lr[i]; g.lif[i] = (float)li[i];
    }
    for (int i = 0; i < kGridMax; ++i) {
        g.dpos_r[i] = (float)(g.lr2[i] - g.lr2[kGridMax - 1]); g.dneg_r[i] = (float)(g.lr2[i] - g.lr2[0]);
        g.dpos_i[i] = (float)(g.li2[i] - g.li2[kGridMax - 1]); g.dneg_i[i] = (float)(g.li2[i] - g.li2[0]);
    }
    // the device code hard-wires the reference table's two irregular grid points (see the kernel)
    bool uniform = true;                                       // equally spaced levels on both axes (the rho^l exponentials)
    for (int i = 2; i < kGridMax; ++i) {
        if (fabs((lr[i] - lr[i - 1]) - (lr[1] - lr[0])) > 1e-12 * fabs(lr[1] - lr[0])) uniform = false;
        if (fabs((li[i] - li[i - 1]) - (li[1] - li[0])) > 1e-12 * fabs(li[1] - li[0])) uniform = false;
    }
    const bool ref16 = uniform && nc == 2 && g.ca[0] == 1 && g.cb[0] == 3 && g.cw[0] == 1.f && g.ca[1] == 2 && g.cb[1] == 0 && g.cw[1] == -1.f;
    g.ok = ref16 ? 1 : 0;
    // corners of the grid in the table (the Loss shortcut of fast_loss2): every corner must be present, with a negative
    // bottom and a positive top level on both axes
    g.corner_ok = (lr[0] < 0 && lr[nr - 1] > 0 && li[0] < 0 && li[ni - 1] > 0) ? 1 : 0;
    g.cmag[0] = (float)-lr[0]; g.cmag[1] = (float)lr[nr - 1]; g.cmag[2] = (float)-li[0]; g.cmag[3] = (float)li[ni - 1];
    for (int sl = 0; sl < 4; ++sl) {
        const double cr = (sl & 2) ? lr[0] : lr[nr - 1], ci = (sl & 1) ? li[0] : li[ni - 1];
        g.corner_k[sl] = -1;
        for (int k = al.K - 1; k >= 0; --k)
            if (al.re[k] == cr && al.im[k] == ci) g.corner_k[sl] = k;
        if (g.corner_k[sl] < 0) g.corner_ok = 0;
    }
    return g;
}

// ---- section reductions over the column-owner layout (column = lane + 32 t) ---------------------------------------
template <int M_, int CP>
__device__ __forceinline__ void section_max(const float (&lmax)[CP], float (&smax)[CP]) {
    if constexpr (M_ >= 32) {
        constexpr int TPS = M_ / 32;
#pragma unroll
        for (int s0 = 0; s0 < CP; s0 += TPS) {
            float m = lmax[s0];
#pragma unroll
            for (int q = 1; q < TPS; ++q) m = fmaxf(m, lmax[s0 + q]);
            m = warp_max(m);
#pragma unroll
            for (int q = 0; q < TPS; ++q) smax[s0 + q] = m;
        }
    } else {
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            float m = lmax[t];
#pragma unroll
            for (int o = M_ / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            smax[t] = m;
        }
    }
}
// Z = section sum of S0; others = Z - S0 WITHOUT cancellation: in a butterfly all-reduce, what a lane receives adds
// up to everybody else's share.
template <int M_, int CP>
__device__ __forceinline__ void section_sum_excl(const float (&S0)[CP], float (&Z)[CP], float (&others)[CP]) {
    if constexpr (M_ >= 32) {
        constexpr int TPS = M_ / 32;
#pragma unroll
        for (int s0 = 0; s0 < CP; s0 += TPS) {
            float mine = S0[s0];
#pragma unroll
            for (int q = 1; q < TPS; ++q) mine += S0[s0 + q];
            float part = mine, recv = 0.f;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float r = __shfl_xor_sync(0xffffffffu, part, o);
                recv += r;
                part += r;
            }
#pragma unroll
            for (int q = 0; q < TPS; ++q) {
                Z[s0 + q] = part;
                float sib = 0.f;                                     // the lane's other columns of this section
#pragma unroll
                for (int w = 0; w < TPS; ++w)
                    if (w != q) sib += S0[s0 + w];
                others[s0 + q] = recv + sib;
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            float part = S0[t], recv = 0.f;
#pragma unroll
            for (int o = M_ / 2; o > 0; o >>= 1) {
                const float r = __shfl_xor_sync(0xffffffffu, part, o);
                recv += r;
                part += r;
            }
            Z[t] = part;
            others[t] = recv;
        }
    }
}
__device__ __forceinline__ float pick4(const float (&v)[4], int i) { return i == 0 ? v[0] : (i == 1 ? v[1] : (i == 2 ? v[2] : v[3])); }

// ---- section denoiser on the column-owner layout (column = lane + 32 t), one warp per frame ---------------------------
// Input: q = s / tau (complex64) of the lane's CP columns; output: posterior mean and variance (bamp.py:66-77,
// vamp.py:96-119).  GRID: the separable path for the reference's 16-QAM table (see bamp_fast.cu); otherwise float64
// exponent products and differences, float32 ex2, the exponentials parked in `ebuf` (shared, 32 * CP * K_ floats)
// between the two passes.  "1 - p" comes from the butterfly's exclusive sum (no cancellation), the variance is the
// reference's two-term form.
template <int N_, int M_, int K_, bool GRID, int CP>
__device__ __forceinline__ void fast_denoise(const float (&q_r)[CP], const float (&q_i)[CP], const DevAlphabet& al, const DevGrid& G,
                                             float* ebuf, int lane, float (&xr_)[CP], float (&xi_)[CP], float (&vn_)[CP]) {
    constexpr int L_ = N_ / M_;
    if constexpr (GRID) {
        // The antenna's largest level product lm = q_r lr_max + q_i li_max is kept as an unevaluated float32 sum hi + lo
        // (products split exactly by FMA, the addition by two-sum), so that the offset to the section maximum -- a
        // difference of two numbers of order |q| -- carries no float32 cancellation error without any float64
        // instruction on the iteration's critical path (the float64 chain it replaces was ~10 dependent conversions,
        // multiplies and adds).  The shift is the float `hi` maximum: any common shift of a section is exact.
        float lmax[CP], smax[CP], lhi[CP], llo[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const float ar = q_r[t] >= 0.f ? G.lr2f[3] : G.lr2f[0], ai = q_i[t] >= 0.f ? G.li2f[3] : G.li2f[0];
            const float ar_lo = q_r[t] >= 0.f ? G.lr2l[3] : G.lr2l[0], ai_lo = q_i[t] >= 0.f ? G.li2l[3] : G.li2l[0];
            const float p1 = q_r[t] * ar, p2 = q_i[t] * ai;
            const float e1 = fmaf(q_r[t], ar, -p1), e2 = fmaf(q_i[t], ai, -p2);
            const float sm = p1 + p2, bb = sm - p1;
            const float err = (p1 - (sm - bb)) + (p2 - bb);
            lhi[t] = sm;
            llo[t] = (e1 + e2) + (err + fmaf(q_r[t], ar_lo, q_i[t] * ai_lo));     // + the float32 rounding of the level constants
            lmax[t] = (lane + 32 * t < N_) ? sm : -INFINITY;
        }
        if constexpr (L_ == 1) {           // the section is the whole warp: one CREDUX instead of a shuffle tree
            float m = lmax[0];
#pragma unroll
            for (int t = 1; t < CP; ++t) m = fmaxf(m, lmax[t]);
            float r;
            asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(m));
#pragma unroll
            for (int t = 0; t < CP; ++t) smax[t] = r;
        } else {
            section_max<M_, CP>(lmax, smax);
        }
        float Er[CP][4], Ei[CP][4], S0[CP], A0[CP], A1[CP], B0[CP], B1[CP], e13[CP], e20[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const float off = (lhi[t] - smax[t]) + llo[t];              // <= 0 up to rounding
            const bool rp = q_r[t] >= 0.f, ip = q_i[t] >= 0.f;
            // equally spaced levels (make_grid checks it): the four exponentials of an axis are 1, rho, rho^2, rho^3 with
            // rho = 2^(-|q| spacing), counted from the antenna's largest level -- 3 MUFU per antenna instead of 8
            const float rho_r = fast_ex2(-fabsf(q_r[t]) * G.dneg_r[1]), rho_i = fast_ex2(-fabsf(q_i[t]) * G.dneg_i[1]);
            const float e0 = fast_ex2(off);
            const float rr2 = rho_r * rho_r, rr3 = rr2 * rho_r;
            const float i1 = e0 * rho_i, i2 = i1 * rho_i, i3 = i2 * rho_i;
            Er[t][0] = rp ? rr3 : 1.0f; Er[t][1] = rp ? rr2 : rho_r; Er[t][2] = rp ? rho_r : rr2; Er[t][3] = rp ? 1.0f : rr3;
            Ei[t][0] = ip ? i3 : e0;    Ei[t][1] = ip ? i2 : i1;     Ei[t][2] = ip ? i1 : i2;     Ei[t][3] = ip ? e0 : i3;
            const float a0 = (Er[t][0] + Er[t][1]) + (Er[t][2] + Er[t][3]);
            const float b0 = (Ei[t][0] + Ei[t][1]) + (Ei[t][2] + Ei[t][3]);
            const float a1 = fmaf(G.lrf[0], Er[t][0], G.lrf[1] * Er[t][1]) + fmaf(G.lrf[2], Er[t][2], G.lrf[3] * Er[t][3]);
            const float b1 = fmaf(G.lif[0], Ei[t][0], G.lif[1] * Ei[t][1]) + fmaf(G.lif[2], Ei[t][2], G.lif[3] * Ei[t][3]);
            e13[t] = Er[t][1] * Ei[t][3];
            e20[t] = Er[t][2] * Ei[t][0];
            const float s0 = fmaf(a0, b0, e13[t] - e20[t]);
            S0[t] = (lane + 32 * t < N_) ? s0 : 0.f;
            A0[t] = a0; A1[t] = a1; B0[t] = b0; B1[t] = b1;
        }
        float Z[CP], others[CP];
        section_sum_excl<M_, CP>(S0, Z, others);
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const float rz = fast_rcp(Z[t]);
            const float s1r = fmaf(A1[t], B0[t], fmaf(G.lrf[1], e13[t], -G.lrf[2] * e20[t]));
            const float s1i = fmaf(A0[t], B1[t], fmaf(G.lif[3], e13[t], -G.lif[0] * e20[t]));
            const float xr = s1r * rz, xi = s1i * rz;
            float er2[4], ei2[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float er = xr - G.lrf[l], ei = xi - G.lif[l];
                er2[l] = er * er;
                ei2[l] = ei * ei;
            }
            const float dr = fmaf(er2[0], Er[t][0], er2[1] * Er[t][1]) + fmaf(er2[2], Er[t][2], er2[3] * Er[t][3]);
            const float di = fmaf(ei2[0], Ei[t][0], ei2[1] * Ei[t][1]) + fmaf(ei2[2], Ei[t][2], ei2[3] * Ei[t][3]);
            float spread = fmaf(dr, B0[t], A0[t] * di);
            spread += fmaf(er2[1] + ei2[3], e13[t], -(er2[2] + ei2[0]) * e20[t]);
            xr_[t] = xr;
            xi_[t] = xi;
            vn_[t] = fmaf(fmaf(xr, xr, xi * xi), others[t] * rz, spread * rz);
        }
    } else {
        double qr[CP], qi[CP];
        float lmax[CP], smax[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            qr[t] = (double)q_r[t];
            qi[t] = (double)q_i[t];
            float m = -INFINITY;
#pragma unroll
            for (int k = 0; k < K_; ++k) m = fmaxf(m, fmaf(q_r[t], al.ref[k], q_i[t] * al.imf[k]));
            lmax[t] = (lane + 32 * t < N_) ? m : -INFINITY;
        }
        section_max<M_, CP>(lmax, smax);      // only approximately the true maxima: a common shift, nothing else
        float S0[CP], S1r[CP], S1i[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const double shift = (double)smax[t];
            float s0 = 0.f, s1r = 0.f, s1i = 0.f;
#pragma unroll
            for (int k = 0; k < K_; ++k) {
                const double x = fma(qr[t], al.re[k], qi[t] * al.im[k]);
                const float e = fast_ex2((float)(x - shift) * 1.4426950408889634f);
                ebuf[(t * K_ + k) * 32 + lane] = e;
                s0 += e;
                s1r = fmaf(al.ref[k], e, s1r);
                s1i = fmaf(al.imf[k], e, s1i);
            }
            S0[t] = (lane + 32 * t < N_) ? s0 : 0.f;
            S1r[t] = s1r;
            S1i[t] = s1i;
        }
        float Z[CP], others[CP];
        section_sum_excl<M_, CP>(S0, Z, others);
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const float rz = fast_rcp(Z[t]);
            const float xr = S1r[t] * rz, xi = S1i[t] * rz;
            float spread = 0.f;
#pragma unroll
            for (int k = 0; k < K_; ++k) {
                const float e = ebuf[(t * K_ + k) * 32 + lane];
                const float dr = xr - al.ref[k], di = xi - al.imf[k];
                spread = fmaf(fmaf(dr, dr, di * di), e, spread);
            }
            xr_[t] = xr;
            xi_[t] = xi;
            vn_[t] = fmaf(fmaf(xr, xr, xi * xi), others[t] * rz, spread * rz);
        }
    }
}

// Every symbol in {0, +-1, +-j} (the reference's OOK / BPSK / QPSK tables, config.py:86-95): the exponent
// q.re s.re + q.im s.im is then ONE exact float32 product, and its float32 difference to the float32 shift is the correctly
// rounded difference -- the same number the float64 path of fast_denoise rounds to float32 before ex2.
inline bool alphabet_is_exact(const DevAlphabet& al) {
    for (int k = 0; k < al.K; ++k) {
        const double ar = al.re[k] < 0 ? -al.re[k] : al.re[k], ai = al.im[k] < 0 ? -al.im[k] : al.im[k];
        if (!((ar == 0.0 || ar == 1.0) && (ai == 0.0 || ai == 1.0) && !(ar == 1.0 && ai == 1.0))) return false;
    }
    return true;
}

// Section denoiser for such alphabets with ONE column per lane (the four-warps-per-frame kernel): same operations in the same
// order as the generic branch of fast_denoise (bit-identical results for these alphabets) without any float64 instruction
// (conversions run on the XU pipe at a sixteenth of the FP32 rate and sat on the iteration's critical path), the exponentials
// in registers, the section maximum by one masked REDUX.
template <int M_, int K_>
__device__ __forceinline__ void exact_denoise1(float q_r, float q_i, const DevAlphabet& al, int lane, float& xr_, float& xi_, float& vn_) {
    static_assert(M_ == 32 || M_ == 16 || M_ == 8, "a section is an aligned group of lanes");
    float x[K_];
    float m = -INFINITY;
#pragma unroll
    for (int k = 0; k < K_; ++k) {
        x[k] = fmaf(q_r, al.ref[k], q_i * al.imf[k]);
        m = fmaxf(m, x[k]);
    }
    float smax;
    if constexpr (M_ == 32) {
        asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(smax) : "f"(m));
    } else {
        const unsigned mask = ((M_ == 16) ? 0xffffu : 0xffu) << (lane & ~(M_ - 1));
        asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(smax) : "f"(m), "r"(mask));
    }
    float e[K_], s0 = 0.f, s1r = 0.f, s1i = 0.f;
#pragma unroll
    for (int k = 0; k < K_; ++k) {
        e[k] = fast_ex2((x[k] - smax) * 1.4426950408889634f);
        s0 += e[k];
        s1r = fmaf(al.ref[k], e[k], s1r);
        s1i = fmaf(al.imf[k], e[k], s1i);
    }
    const float S0[1] = {s0};
    float Z[1], others[1];
    section_sum_excl<M_, 1>(S0, Z, others);
    const float rz = fast_rcp(Z[0]);
    const float xr = s1r * rz, xi = s1i * rz;
    float spread = 0.f;
#pragma unroll
    for (int k = 0; k < K_; ++k) {
        const float dr = xr - al.ref[k], di = xi - al.imf[k];
        spread = fmaf(fmaf(dr, dr, di * di), e[k], spread);
    }
    xr_ = xr;
    xi_ = xi;
    vn_ = fmaf(fmaf(xr, xr, xi * xi), others[0] * rz, spread * rz);
}

// ---- Loss on the column-owner layout: MAP decision + counters (loss.py:282-302, 67-179), Lin = 1, one warp per frame ----
// The first version cost ~3600 cycles per frame on B200 (phase clocks, scripts/phase_clocks.py) -- more than one whole
// BAMP iteration -- almost all of it exposed latency: a 15-step float64 compare chain per column, a 5-step shuffle
// butterfly on (double, int) picks, x_true / label loads from L2 on first use, and two more 5-step butterflies for the
// counters.  Here (a) the inputs are staged into shared memory by cp.async at the START of the frame, (b) the per-column
// arg-max is a tournament (depth log2 K), (c) the section arg-max is three REDUX on an order-preserving 64-bit key,
// (d) counters accumulate per lane / by fire-and-forget shared atomics and are folded once per kernel.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst_smem, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
template <int PENDING>
__device__ __forceinline__ void cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// shared-memory staging of one frame's Loss inputs: x_true [N] float2 | idx_true [L] int64 | sym_true [L] int64
template <int N_, int L_>
struct LossStage {
    static constexpr int lab_bytes = (L_ * 8 + 15) & ~15;
    static constexpr int bytes = N_ * 8 + 2 * lab_bytes;
    __device__ static const float2* xt(const unsigned char* st) { return reinterpret_cast<const float2*>(st); }
    __device__ static const long long* idx(const unsigned char* st) { return reinterpret_cast<const long long*>(st + N_ * 8); }
    __device__ static const long long* sym(const unsigned char* st) { return reinterpret_cast<const long long*>(st + N_ * 8 + lab_bytes); }
    // x_true + f N must be 16-byte aligned (the launchers check the base pointer; N is even for every fast shape)
    // `lane` = index of the calling thread among the `nthreads` that share the stage (one warp, or the CTA of the
    // four-warps-per-frame kernel); every one of them commits exactly one cp.async group
    __device__ static void issue(unsigned char* st, const LossIO& io, long long f, int lane, int nthreads = 32) {
        static_assert(N_ % 2 == 0, "x_true rows are copied in 16-byte pieces");
        for (int c = lane; c < N_ / 2; c += nthreads) cp_async16(st + c * 16, io.x_true + f * N_ + 2 * c);
        if (lane < L_) {
            cp_async8(st + N_ * 8 + lane * 8, io.idx_true + f * L_ + lane);
            cp_async8(st + N_ * 8 + lab_bytes + lane * 8, io.sym_true + f * L_ + lane);
        }
        cp_async_commit();
    }
};

// order-preserving key of np.argmax: NaN beats everything, -0.0 == +0.0
__device__ __forceinline__ unsigned long long argmax_key(double v) {
    const long long b = __double_as_longlong(__dadd_rn(v, 0.0));          // -0.0 + 0.0 = +0.0
    const unsigned long long k = b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
    return (v != v) ? ~0ull : k;
}
// first maximum of `key` over the lanes of a section (M_ lanes, or the whole warp), ties to the smallest idx
template <int M_>
__device__ __forceinline__ int section_argmax(unsigned long long key, int idx, int lane) {
    unsigned mask = 0xffffffffu;
    if constexpr (M_ < 32) mask = ((1u << M_) - 1u) << (lane & ~(M_ - 1));
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mh = __reduce_max_sync(mask, hi);
    const unsigned ml = __reduce_max_sync(mask, hi == mh ? lo : 0u);
    return (int)__reduce_min_sync(mask, (hi == mh && lo == ml) ? (unsigned)idx : 0x7fffffffu);
}

// xmap / xh: the lane's CP columns (column = lane + 32 t) of the decision input and of the MMSE estimate; `st` the staged
// inputs of this frame (LossStage, complete: cp.async.wait_all + __syncwarp done by the caller).  cnt32: the warp's 16
// shared counters (Counter enum slots), sqacc: 32 per-lane float64 squared-error sums.
// GRID (product-grid alphabets, make_grid): Re(x conj(s)) = x.re s.re + x.im s.im is maximised by the corner of the grid
// that carries the signs of x, by a margin of (level spacing) x min(|x.re|, |x.im|) before rounding; whenever that margin
// exceeds the float64 rounding of the two products and the sum by orders of magnitude (min > 1e-12 max, both finite) the
// corner is the UNIQUE maximum of the rounded values as well, so np.argmax's answer for the column is the corner's first
// table index and only its metric needs to be evaluated (3 float64 instructions instead of ~50 + a 15-merge tournament).
// Columns that fail the test (zeros, NaN, Inf, extreme ratios) take the full tournament.
// col_base: the frame column of lane 0's first column (0 for the one-warp-per-frame kernels; 32 w for warp w of the
// four-warps-per-frame kernel, whose sections must then lie inside one warp).  Returns bit 0: some decided value of this
// warp's columns differs from x_true, bit 1: NaN input; with book_frame the frame-level counters are booked here.
template <int N_, int M_, int K_, int CP, bool GRID>
__device__ __forceinline__ unsigned fast_loss2(const float2 (&xmap)[CP], const float2 (&xh)[CP], const DevAlphabet& al, const DevGrid& G,
                                               const Geom& g, const unsigned char* st, long long f, int lane, unsigned* cnt32,
                                               double* sqacc, int col_base = 0, bool book_frame = true) {
    constexpr int L_ = N_ / M_;
    using LS = LossStage<N_, L_>;
    bool nan_seen = false;
    int dec_ant[CP], dec_k[CP];
    bool fast = false;
    // Float32 shortcut for a frame that is ONE section of a product-grid alphabet (BASELINE config 2): every column's best
    // symbol is its corner (see above) and its metric is |x.re| |level_re| + |x.im| |level_im|, which float32 evaluates to 2e-7
    // relative.  If exactly one column lies within 1e-6 (relative) of the float32 maximum, it is the strict maximum of the
    // float64 metrics as well, so np.argmax's answer is known without any float64 instruction; otherwise (near-ties, NaN,
    // Inf, zeros) the exact path below decides.  All conditions are warp-uniform.
    if constexpr (GRID && L_ == 1 && M_ >= 32) {
        bool ok = G.corner_ok != 0;
        float w[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            const float ax = fabsf(xmap[t].x), ay = fabsf(xmap[t].y);
            const bool live = col_base + lane + 32 * t < N_;
            ok &= !live || (fminf(ax, ay) > 1e-12f * fmaxf(ax, ay) && fmaxf(ax, ay) < INFINITY);      // false for NaN
            w[t] = live ? fmaf(ax, xmap[t].x < 0.f ? G.cmag[0] : G.cmag[1], ay * (xmap[t].y < 0.f ? G.cmag[2] : G.cmag[3])) : 0.f;
        }
        if (__all_sync(0xffffffffu, ok)) {
            float wl = w[0];
#pragma unroll
            for (int t = 1; t < CP; ++t) wl = fmaxf(wl, w[t]);
            const float wmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(wl)));      // w >= 0: bit order = value order
            const float thr = wmax * (1.0f - 1.0e-6f);
            int cand = 0, tc = 0;
#pragma unroll
            for (int t = 0; t < CP; ++t) {
                const bool c = w[t] >= thr;
                cand += c;
                tc = c ? t : tc;
            }
            const unsigned holders = __ballot_sync(0xffffffffu, cand > 0);
            if (__popc(holders) == 1 && __reduce_add_sync(0xffffffffu, (unsigned)cand) == 1u) {
                const int wlane = __ffs(holders) - 1;
                float2 xw = xmap[0];
#pragma unroll
                for (int t = 1; t < CP; ++t) xw = (tc == t) ? xmap[t] : xw;
                int mine = (col_base + lane + 32 * tc) % M_ * K_ + G.corner_k[(xw.x < 0.f ? 2 : 0) + (xw.y < 0.f ? 1 : 0)];
                mine = __shfl_sync(0xffffffffu, mine, wlane);
#pragma unroll
                for (int t = 0; t < CP; ++t) {
                    dec_ant[t] = mine / K_;
                    dec_k[t] = mine % K_;
                }
                fast = true;
            }
        }
    }
    if (!fast) {
        unsigned long long key[CP];
        int kidx[CP];
    #pragma unroll
        for (int t = 0; t < CP; ++t) {
            const int col = col_base + lane + 32 * t;
            key[t] = 0ull;
            kidx[t] = 0x7fffffff;
            if (col < N_) {
                // Re(x conj(sym_k)) in complex128 for every k (loss.py:295), then a tournament over ordered neighbours: the
                // right-hand winner replaces the left-hand one only if it is strictly greater or NaN and the left is not NaN
                // -- np.argmax's first-maximum / first-NaN rule for any merge of two index-ordered groups.
                const double xr = (double)xmap[t].x, xi = (double)xmap[t].y;
                bool corner = false;
                if constexpr (GRID) {
                    const float ax = fabsf(xmap[t].x), ay = fabsf(xmap[t].y);
                    corner = G.corner_ok && fminf(ax, ay) > 1e-12f * fmaxf(ax, ay) && fmaxf(ax, ay) < INFINITY;   // false for NaN
                }
                double wv;
                int wk;
                if (corner) {
                    wk = G.corner_k[(xmap[t].x < 0.f ? 2 : 0) + (xmap[t].y < 0.f ? 1 : 0)];
                    wv = __dadd_rn(__dmul_rn(xr, al.re[wk]), __dmul_rn(xi, al.im[wk]));
                } else {
                    double bv[K_];
                    int bk[K_];
    #pragma unroll
                    for (int k = 0; k < K_; ++k) {
                        bv[k] = __dadd_rn(__dmul_rn(xr, al.re[k]), __dmul_rn(xi, al.im[k]));
                        bk[k] = k;
                    }
    #pragma unroll
                    for (int s = 1; s < K_; s *= 2) {
    #pragma unroll
                        for (int i = 0; i + s < K_; i += 2 * s) {
                            const bool upd = (bv[i] == bv[i]) & ((bv[i + s] > bv[i]) | (bv[i + s] != bv[i + s]));
                            bv[i] = upd ? bv[i + s] : bv[i];
                            bk[i] = upd ? bk[i + s] : bk[i];
                        }
                    }
                    wv = bv[0];
                    wk = bk[0];
                }
                key[t] = argmax_key(wv);
                kidx[t] = (col % M_) * K_ + wk;
                nan_seen |= (xmap[t].x != xmap[t].x) || (xmap[t].y != xmap[t].y);
            }
        }
        if constexpr (M_ >= 32) {
            constexpr int TPS = M_ / 32;
    #pragma unroll
            for (int s0 = 0; s0 < CP; s0 += TPS) {
                unsigned long long bkey = key[s0];
                int bidx = kidx[s0];
    #pragma unroll
                for (int q = 1; q < TPS; ++q) {                       // the lane's later columns have larger flat indices
                    const bool upd = key[s0 + q] > bkey;
                    bkey = upd ? key[s0 + q] : bkey;
                    bidx = upd ? kidx[s0 + q] : bidx;
                }
                const int w = section_argmax<32>(bkey, bidx, lane);
    #pragma unroll
                for (int q = 0; q < TPS; ++q) {
                    dec_ant[s0 + q] = w / K_;
                    dec_k[s0 + q] = w % K_;
                }
            }
        } else {
    #pragma unroll
            for (int t = 0; t < CP; ++t) {
                const int w = section_argmax<M_>(key[t], kidx[t], lane);
                dec_ant[t] = w / K_;
                dec_k[t] = w % K_;
            }
        }
    }
    bool wrong = false;
    double sq = 0.0;
#pragma unroll
    for (int t = 0; t < CP; ++t) {
        const int col = col_base + lane + 32 * t;
        if (col < N_) {
            const int sec = col / M_, m = col % M_;
            const float2 xt = LS::xt(st)[col];
            const int k = dec_k[t];
            const float2 h = (m == dec_ant[t]) ? make_float2((float)al.re[k], (float)al.im[k]) : make_float2(0.f, 0.f);
            wrong |= (h.x != xt.x) || (h.y != xt.y);
            const float dr = xh[t].x - xt.x, di = xh[t].y - xt.y;
            sq += (double)dr * dr + (double)di * di;
            if (m == 0) {   // one lane per section books the label counters
                const long long ih = (g.frame_base + f) * (long long)N_ + sec * M_ + dec_ant[t];
                const long long itrue = LS::idx(st)[sec];
                const long long sh = al.gray[k], strue = LS::sym(st)[sec];
                const unsigned long long imask = g.index_bits_kept >= 64 ? ~0ull : ((1ull << g.index_bits_kept) - 1ull);
                const unsigned ib = __popcll((unsigned long long)(ih ^ itrue) & imask);
                const unsigned sb = __popcll((unsigned long long)(sh ^ strue) & ((1ull << al.sbits) - 1ull));
                if (ih != itrue) atomicAdd(&cnt32[C_INDEX_ERR], 1u);
                if (sh != strue) atomicAdd(&cnt32[C_SYMBOL_ERR], 1u);
                if (ib) atomicAdd(&cnt32[C_INDEX_BIT], ib);
                if (sb) atomicAdd(&cnt32[C_SYMBOL_BIT], sb);
            }
        }
    }
    sqacc[lane] += sq;
    const bool any_wrong = __any_sync(0xffffffffu, wrong), any_nan = __any_sync(0xffffffffu, nan_seen);
    if (book_frame && lane == 0) {
        if (any_wrong) atomicAdd(&cnt32[C_FRAME_ERR], 1u);             // Lin = 1: one time slot per frame
        if (any_nan) atomicAdd(&cnt32[C_NAN_FRAMES], 1u);
    }
    return (any_wrong ? 1u : 0u) | (any_nan ? 2u : 0u);
}
// fold a warp's shared counters into the global block (Lin = 1: the frame is its only, first, middle and last slot).
// 32-bit per-warp counts: a warp would need > 6e7 frames of 64 label bits in one call to overflow.
__device__ __forceinline__ void fast_flush2(const unsigned* cnt32, const double* sqacc, unsigned long long* out, int lane) {
    __syncwarp();
    const double sq = warp_sum(sqacc[lane]);
    if (lane != 0 || !out) return;
    const int plain[] = {C_FRAMES, C_INDEX_ERR, C_SYMBOL_ERR, C_INDEX_BIT, C_SYMBOL_BIT, C_ITERS, C_NAN_FRAMES};
    for (int k : plain)
        if (cnt32[k]) atomicAdd(out + k, (unsigned long long)cnt32[k]);
    if (cnt32[C_FRAME_ERR]) {
        const int slots[] = {C_FRAME_ERR, C_SLOT_ERR, C_SLOT_FIRST, C_SLOT_MID, C_SLOT_LAST};
        for (int k : slots) atomicAdd(out + k, (unsigned long long)cnt32[C_FRAME_ERR]);
    }
    if (sq != 0.0)
        for (int k = 0; k < 4; ++k) atomicAdd(reinterpret_cast<double*>(out) + C_SQERR + k, sq);
}

}  // namespace ampsm
