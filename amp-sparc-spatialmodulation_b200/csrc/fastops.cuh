// Device helpers shared by the register-resident kernels (bamp_fast.cu, bamp_pair.cu, vamp_pair.cu): packed
// fp32x2 arithmetic, approximate MUFU wrappers, the product-grid view of an alphabet and the section reductions
// over the column-owner layout.
#pragma once
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
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
        g.lrf[i] = (float)lr[i]; g.lif[i] = (float)li[i];
    }
    for (int i = 0; i < kGridMax; ++i) {
        g.dpos_r[i] = (float)(g.lr2[i] - g.lr2[kGridMax - 1]); g.dneg_r[i] = (float)(g.lr2[i] - g.lr2[0]);
        g.dpos_i[i] = (float)(g.li2[i] - g.li2[kGridMax - 1]); g.dneg_i[i] = (float)(g.li2[i] - g.li2[0]);
    }
    // the device code hard-wires the reference table's two irregular grid points (see the kernel)
    const bool ref16 = nc == 2 && g.ca[0] == 1 && g.cb[0] == 3 && g.cw[0] == 1.f && g.ca[1] == 2 && g.cb[1] == 0 && g.cw[1] == -1.f;
    g.ok = ref16 ? 1 : 0;
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

}  // namespace ampsm
