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
        float lmax[CP], smax[CP];
        double lmd[CP];
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            lmd[t] = (double)q_r[t] * (q_r[t] >= 0.f ? G.lr2[3] : G.lr2[0]) + (double)q_i[t] * (q_i[t] >= 0.f ? G.li2[3] : G.li2[0]);
            lmax[t] = (lane + 32 * t < N_) ? (float)lmd[t] : -INFINITY;
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
            const float off = (float)(lmd[t] - (double)smax[t]);       // <= 0 up to rounding
            const bool rp = q_r[t] >= 0.f, ip = q_i[t] >= 0.f;
            float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                Er[t][l] = fast_ex2(q_r[t] * (rp ? G.dpos_r[l] : G.dneg_r[l]));
                Ei[t][l] = fast_ex2(fmaf(q_i[t], ip ? G.dpos_i[l] : G.dneg_i[l], off));
                a0 += Er[t][l];
                a1 = fmaf(G.lrf[l], Er[t][l], a1);
                b0 += Ei[t][l];
                b1 = fmaf(G.lif[l], Ei[t][l], b1);
            }
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
            float dr = 0.f, di = 0.f, er2[4], ei2[4];
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                const float er = xr - G.lrf[l], ei = xi - G.lif[l];
                er2[l] = er * er;
                ei2[l] = ei * ei;
                dr = fmaf(er2[l], Er[t][l], dr);
                di = fmaf(ei2[l], Ei[t][l], di);
            }
            float spread = fmaf(dr, B0[t], A0[t] * di);
            spread = fmaf(er2[1] + ei2[3], e13[t], spread);
            spread = fmaf(-(er2[2] + ei2[0]), e20[t], spread);
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

// ---- Loss on the column-owner layout: MAP decision + counters (loss.py:282-302, 67-179), Lin = 1, one warp per frame ----
// xmap / xh: the lane's CP columns of the decision input and of the MMSE estimate.  Books into the warp's shared
// counter block `cnt` (slots as the Counter enum, slot 12 = squared-error sum as double); lane 0 writes.
template <int N_, int M_, int K_, int CP>
__device__ __forceinline__ void fast_loss(const float2 (&xmap)[CP], const float2 (&xh)[CP], const DevAlphabet& al, const Geom& g,
                                          const LossIO& io, long long f, int lane, unsigned long long* cnt) {
    constexpr int L_ = N_ / M_;
    bool wrong = false, nan_seen = false;
    double sq = 0.0;
    Pick best[CP];
#pragma unroll
    for (int t = 0; t < CP; ++t) {
        const int col = lane + 32 * t;
        best[t] = Pick{-INFINITY, 0x7fffffff};
        if (col < N_) {
            const int m = col % M_;
            // in-order scan (flat index increases with k): the first maximum wins, so replace only on "strictly
            // greater"; a NaN wins once and then sticks (np.argmax).  Predicated selects, no branches.
            const double xr = (double)xmap[t].x, xi = (double)xmap[t].y;
            double bv = __dadd_rn(__dmul_rn(xr, al.re[0]), __dmul_rn(xi, al.im[0]));
            int bk = 0;
#pragma unroll
            for (int k = 1; k < K_; ++k) {
                const double v = __dadd_rn(__dmul_rn(xr, al.re[k]), __dmul_rn(xi, al.im[k]));
                const bool upd = (bv == bv) & ((v > bv) | (v != v));
                bv = upd ? v : bv;
                bk = upd ? k : bk;
            }
            best[t] = Pick{bv, m * K_ + bk};
            nan_seen |= (xmap[t].x != xmap[t].x) || (xmap[t].y != xmap[t].y);
        }
    }
    int dec_ant[CP], dec_k[CP];
    if constexpr (M_ >= 32) {
        constexpr int TPS = M_ / 32;
#pragma unroll
        for (int s0 = 0; s0 < CP; s0 += TPS) {
            Pick b = best[s0];
#pragma unroll
            for (int q = 1; q < TPS; ++q)
                if (pick_better(best[s0 + q], b)) b = best[s0 + q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                Pick other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.idx, o)};
                if (pick_better(other, b)) b = other;
            }
#pragma unroll
            for (int q = 0; q < TPS; ++q) {
                dec_ant[s0 + q] = b.idx / K_;
                dec_k[s0 + q] = b.idx % K_;
            }
        }
    } else {
#pragma unroll
        for (int t = 0; t < CP; ++t) {
            Pick b = best[t];
#pragma unroll
            for (int o = M_ / 2; o > 0; o >>= 1) {
                Pick other{__shfl_xor_sync(0xffffffffu, b.v, o), __shfl_xor_sync(0xffffffffu, b.idx, o)};
                if (pick_better(other, b)) b = other;
            }
            dec_ant[t] = b.idx / K_;
            dec_k[t] = b.idx % K_;
        }
    }
    unsigned long long c_idx = 0, c_sym = 0, c_ibit = 0, c_sbit = 0;
#pragma unroll
    for (int t = 0; t < CP; ++t) {
        const int col = lane + 32 * t;
        if (col < N_) {
            const int sec = col / M_, m = col % M_;
            const float2 xt = io.x_true[f * N_ + col];
            const int k = dec_k[t];
            const float2 h = (m == dec_ant[t]) ? make_float2((float)al.re[k], (float)al.im[k]) : make_float2(0.f, 0.f);
            wrong |= (h.x != xt.x) || (h.y != xt.y);
            const float dr = xh[t].x - xt.x, di = xh[t].y - xt.y;
            sq += (double)dr * dr + (double)di * di;
            if (m == 0) {   // one lane per section books the label counters
                const long long ih = (g.frame_base + f) * (long long)N_ + sec * M_ + dec_ant[t];
                const long long itrue = io.idx_true[f * L_ + sec];
                const long long sh = al.gray[k], st = io.sym_true[f * L_ + sec];
                const unsigned long long imask = g.index_bits_kept >= 64 ? ~0ull : ((1ull << g.index_bits_kept) - 1ull);
                c_idx += (ih != itrue);
                c_sym += (sh != st);
                c_ibit += __popcll((unsigned long long)(ih ^ itrue) & imask);
                c_sbit += __popcll((unsigned long long)(sh ^ st) & ((1ull << al.sbits) - 1ull));
            }
        }
    }
    unsigned long long packed = c_idx | (c_sym << 12) | (c_ibit << 24) | (c_sbit << 44);   // <= 64 sections, 64 bits each
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
    sq = warp_sum(sq);
    const bool any_wrong = __any_sync(0xffffffffu, wrong), any_nan = __any_sync(0xffffffffu, nan_seen);
    if (lane == 0) {
        cnt[C_INDEX_ERR] += packed & 0xfffull;
        cnt[C_SYMBOL_ERR] += (packed >> 12) & 0xfffull;
        cnt[C_INDEX_BIT] += (packed >> 24) & 0xfffffull;
        cnt[C_SYMBOL_BIT] += packed >> 44;
        cnt[C_FRAME_ERR] += any_wrong;                       // Lin = 1: one time slot per frame
        cnt[C_NAN_FRAMES] += any_nan;
        reinterpret_cast<double*>(cnt)[12] += sq;
    }
}
// flush a warp's shared counter block into the global one (Lin = 1: the frame is its only, first, middle and last slot)
__device__ __forceinline__ void fast_flush_counters(const unsigned long long* cnt, unsigned long long* out) {
    const int plain[] = {C_FRAMES, C_INDEX_ERR, C_SYMBOL_ERR, C_INDEX_BIT, C_SYMBOL_BIT, C_ITERS, C_NAN_FRAMES};
    for (int k : plain)
        if (cnt[k]) atomicAdd(out + k, cnt[k]);
    if (cnt[C_FRAME_ERR]) {
        const int slots[] = {C_FRAME_ERR, C_SLOT_ERR, C_SLOT_FIRST, C_SLOT_MID, C_SLOT_LAST};
        for (int k : slots) atomicAdd(out + k, cnt[C_FRAME_ERR]);
    }
    const double sq = reinterpret_cast<const double*>(cnt)[12];
    if (sq != 0.0)
        for (int k = 0; k < 4; ++k) atomicAdd(reinterpret_cast<double*>(out) + C_SQERR + k, sq);
}

}  // namespace ampsm
