// Block-cooperative building blocks of the generic (shared-memory resident) kernels:
//   * section-wise spatial-modulation denoiser  (bamp.py:66-77, vamp.py:96-119, scamp.py:61-68)
//   * hard decision + error counters            (loss.py:67-179, 223-250, 282-302)
// One warp owns one section at a time; every reduction is a warp shuffle, so the only block barriers are the
// ones the callers place between phases.
#pragma once
#include "common.cuh"

namespace ampsm {

template <bool EXP64>
struct ExpT;
template <>
struct ExpT<true> {
    using type = double;
};
template <>
struct ExpT<false> {
    using type = float;
};

// Exponent x_mk = Re(q_m conj(sym_k)).  float64 path: products of the complex64-rounded q with the complex128
// symbols, as the reference evaluates them (bamp.py:69).
template <bool EXP64, typename CT>
__device__ __forceinline__ typename ExpT<EXP64>::type sm_exponent(CT q, const DevAlphabet& al, int k) {
    if constexpr (EXP64) {
        return __dadd_rn(__dmul_rn((double)q.x, al.re[k]), __dmul_rn((double)q.y, al.im[k]));
    } else {
        return fmaf((float)q.x, al.ref[k], (float)q.y * al.imf[k]);
    }
}
// complex / real with the operand types of the caller (complex64/float32 or complex128/float64)
__device__ __forceinline__ double2 cdiv_real(double2 a, double d) {
    const double r = __drcp_rn(d);
    return make_double2(__dmul_rn(a.x, r), __dmul_rn(a.y, r));
}
template <typename CT>
struct RealOf;
template <>
struct RealOf<float2> {
    using type = float;
};
template <>
struct RealOf<double2> {
    using type = double;
};
// exp(x - ref): float64 exp as the reference, or -- the fast mode -- a float32 exp of the float64 difference.  The
// exponents reach |x| ~ 1e3 at high SNR, so x and x - ref are always formed in float64: a float32 product
// would carry ~|x| 2^-24 absolute error into the posterior (5e-4 relative on the C1 fixture at 20 dB).
template <bool EXP64>
__device__ __forceinline__ typename ExpT<EXP64>::type exp_shifted(double x, double ref) {
    if constexpr (EXP64) {
        return exp(x - ref);
    } else {
        return __expf((float)(x - ref));
    }
}

// Frame-global max |x| over all (antenna, symbol) entries in float64 -- the reference's shift (bamp.py:70).
// Block-cooperative; `red` is a shared scratch of >= 32 doubles.  Returns the same value in every thread.
template <typename CT>
__device__ inline double block_absmax_exponent(const Geom& g, const DevAlphabet& al, const CT* s,
                                               const typename RealOf<CT>::type* tau_vec,
                                               typename RealOf<CT>::type tau_scalar, bool halve, double* red,
                                               int tau_div = 1) {
    using RT = typename RealOf<CT>::type;
    double m = 0.0;
    for (int j = threadIdx.x; j < g.N; j += blockDim.x) {
        RT tau = tau_vec ? tau_vec[j / tau_div] : tau_scalar;
        if (halve) tau = tau / (RT)2;
        CT q = cdiv_real(s[j], tau);
        for (int k = 0; k < al.K; ++k) m = nanmax(m, fabs(sm_exponent<true>(q, al, k)));
    }
    m = warp_nanmax(m);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r = nanmax(r, red[w]);
    __syncthreads();
    return r;
}

// Denoiser scratch: per-entry planes (3 N) unless a per-warp section buffer is smaller (frames with many short sections)
constexpr int kDenoiseMaxWarps = 16;
__host__ __device__ inline bool denoise_scratch_per_warp(const Geom& g) { return (long long)kDenoiseMaxWarps * g.M < g.N; }
__host__ __device__ inline size_t denoise_scratch_elems(const Geom& g) {
    return 3 * (size_t)(denoise_scratch_per_warp(g) ? kDenoiseMaxWarps * g.M : g.N);
}

// Section-wise posterior mean / variance.  s, tau_vec, xh_out, var_out and the scratch arrays live in shared
// memory (scratch: 3*N values of the exponent type).  No block barrier inside; callers synchronise before
// reading xh_out / var_out.  var_out may be nullptr (SCAMP needs the mean only).
template <bool EXP64, typename CT>
__device__ inline void block_denoise(const Geom& g, const DevAlphabet& al, const CT* s,
                                     const typename RealOf<CT>::type* tau_vec, typename RealOf<CT>::type tau_scalar,
                                     bool halve, double global_shift, float2* xh_out, float* var_out,
                                     typename ExpT<EXP64>::type* scr, int tau_div = 1, bool scr_per_warp = false,
                                     bool var_second_moment = false) {
    using E = typename ExpT<EXP64>::type;
    using RT = typename RealOf<CT>::type;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    // scratch: three planes indexed by the entry (3 N values), or -- scr_per_warp -- by (warp, position in the section):
    // a warp only ever needs its current section between the passes (3 * kDenoiseMaxWarps * M values, see denoise_scratch_elems)
    const int plane = scr_per_warp ? kDenoiseMaxWarps * g.M : g.N;
    E* S0 = scr;
    E* S1r = scr + plane;
    E* S1i = scr + 2 * plane;
    const bool ref_shift = EXP64 && g.shift_mode == 1;
    for (int sec = warp; sec < g.L; sec += nwarps) {
        const int base = sec * g.M;
        const int sb = scr_per_warp ? warp * g.M : base;
        // pass 1: section maximum of the exponents
        double smax = -INFINITY;
        if (!ref_shift) {
            for (int m = lane; m < g.M; m += 32) {
                RT tau = tau_vec ? tau_vec[(base + m) / tau_div] : tau_scalar;
                if (halve) tau = tau / (RT)2;
                CT q = cdiv_real(s[base + m], tau);
#pragma unroll 4
                for (int k = 0; k < al.K; ++k) {
                    const double x = sm_exponent<true>(q, al, k);
                    smax = (x > smax || x != x) ? x : smax;   // NaN sticks
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double other = __shfl_xor_sync(0xffffffffu, smax, o);
                smax = (smax != smax) ? smax : ((other != other || other > smax) ? other : smax);
            }
        } else {
            smax = global_shift;
        }
        // pass 2: per-antenna partial sums
        double z_lane = 0.0;
        for (int m = lane; m < g.M; m += 32) {
            RT tau = tau_vec ? tau_vec[(base + m) / tau_div] : tau_scalar;
            if (halve) tau = tau / (RT)2;
            CT q = cdiv_real(s[base + m], tau);
            E s0 = 0, s1r = 0, s1i = 0;
#pragma unroll 4
            for (int k = 0; k < al.K; ++k) {       // independent exponentials: unrolled so that their latency chains overlap
                E e = exp_shifted<EXP64>(sm_exponent<true>(q, al, k), smax);
                s0 += e;
                if constexpr (EXP64) {
                    s1r += al.re[k] * e;
                    s1i += al.im[k] * e;
                } else {
                    s1r = fmaf(al.ref[k], e, s1r);
                    s1i = fmaf(al.imf[k], e, s1i);
                }
            }
            S0[sb + m] = s0;
            S1r[sb + m] = s1r;
            S1i[sb + m] = s1i;
            z_lane += (double)s0;
        }
        const double Z = warp_sum(z_lane);
        __syncwarp();
        // pass 3: mean, and the variance in the reference's two-term form (bamp.py:74-76)
        for (int m = lane; m < g.M; m += 32) {
            const double xr = (double)S1r[sb + m] / Z, xi = (double)S1i[sb + m] / Z;
            xh_out[base + m] = make_float2((float)xr, (float)xi);
            if (var_out) {
                RT tau = tau_vec ? tau_vec[(base + m) / tau_div] : tau_scalar;
                if (halve) tau = tau / (RT)2;
                CT q = cdiv_real(s[base + m], tau);
                if (var_second_moment) {       // vamp2.py:84-87: E|s|^2 - |E s|^2, as written there (can round below zero)
                    double m2 = 0.0;
#pragma unroll 4
                    for (int k = 0; k < al.K; ++k) {
                        E e = exp_shifted<EXP64>(sm_exponent<true>(q, al, k), smax);
                        const double mag = hypot(al.re[k], al.im[k]);
                        m2 += mag * mag * (double)e;
                    }
                    const double xa = hypot(xr, xi);
                    var_out[base + m] = (float)(m2 / Z - xa * xa);
                    continue;
                }
                double spread = 0.0;
#pragma unroll 4
                for (int k = 0; k < al.K; ++k) {
                    E e = exp_shifted<EXP64>(sm_exponent<true>(q, al, k), smax);
                    const double dr = xr - al.re[k], di = xi - al.im[k];
                    spread += (dr * dr + di * di) * (double)e;
                }
                const double p = (double)S0[sb + m] / Z;
                var_out[base + m] = (float)((xr * xr + xi * xi) * (1.0 - p) + spread / Z);
            }
        }
        __syncwarp();
    }
}

// ``random`` mode (bamp.py:79-101): i.i.d. prior P0 delta_0 + Ps sum_k delta_{s_k}; G(s) = exp(-|r - s|^2 / cov) in float64 with
// no shift, an exactly-zero normaliser replaced by 1e-9.  Ps / P0 are float32 tensors in the reference (bamp.py:40).
__device__ inline void block_denoise_iid(const Geom& g, const DevAlphabet& al, const float2* r, const float* cov, float2* xh_out,
                                         float* var_out) {
    const double sparsity = (double)g.Na / (double)g.Nt;
    const double Ps = (double)(float)(sparsity / al.K), P0 = (double)(float)(1.0 - sparsity);
    for (int j = threadIdx.x; j < g.N; j += blockDim.x) {
        const double rr = (double)r[j].x, ri = (double)r[j].y, c = (double)cov[j];
        const double a0 = hypot(rr, ri);
        double norm = P0 * exp(-(a0 * a0) / c);
        double s1r = 0.0, s1i = 0.0, s2 = 0.0, gs = 0.0;
        for (int k = 0; k < al.K; ++k) {
            const double d = hypot(rr - al.re[k], ri - al.im[k]);
            const double G = exp(-(d * d) / c);
            const double m = hypot(al.re[k], al.im[k]);
            gs += G;
            s1r += al.re[k] * G;
            s1i += al.im[k] * G;
            s2 += m * m * G;
        }
        norm += Ps * gs;
        if (norm == 0.0) norm = 1e-9;                                  // regularize_zero (bamp.py:99-101)
        const double er = Ps * s1r / norm, ei = Ps * s1i / norm;
        const double ea = hypot(er, ei);
        xh_out[j] = make_float2((float)er, (float)ei);
        var_out[j] = (float)(Ps * s2 / norm - ea * ea);
    }
}

// ---- hard decision ------------------------------------------------------------------------------------------
struct Pick {
    double v;
    int idx;
};
// np.argmax order: NaN beats everything, then larger value, then smaller flat index (loss.py:296)
// (written with bitwise predicate logic so that it compiles to SETP/SEL, not branches)
__device__ __forceinline__ bool pick_better(const Pick& a, const Pick& b) {
    const bool an = a.v != a.v, bn = b.v != b.v, first = a.idx < b.idx;
    return (an & (!bn | first)) | (!an & !bn & ((a.v > b.v) | ((a.v == b.v) & first)));
}

// Decide one section with one warp.  Returns (antenna, symbol index) in every lane; symbol index < 0 when the
// segmented rule finds no finite distance (the reference then leaves the section empty).
template <typename CT>
__device__ inline void warp_decide(const Geom& g, const DevAlphabet& al, const CT* xmap_sec, int& ant, int& sym) {
    const int lane = threadIdx.x & 31;
    if (g.decision == 0) {
        // MAP: first maximum of Re(x_m conj(sym_k)) over row-major (m, k), complex128 arithmetic (loss.py:295-296)
        Pick best{-INFINITY, 0x7fffffff};
        const int total = g.M * al.K;
        for (int e = lane; e < total; e += 32) {
            const int m = e / al.K, k = e - m * al.K;
            const CT x = xmap_sec[m];
            Pick c{__dadd_rn(__dmul_rn((double)x.x, al.re[k]), __dmul_rn((double)x.y, al.im[k])), e};
            if (pick_better(c, best)) best = c;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Pick other{__shfl_xor_sync(0xffffffffu, best.v, o), __shfl_xor_sync(0xffffffffu, best.idx, o)};
            if (pick_better(other, best)) best = other;
        }
        ant = best.idx / al.K;
        sym = best.idx - ant * al.K;
    } else {
        // segmented: antenna of largest |x| (argsort()[-1]: NaN sorts last, ties -> highest index), then the
        // nearest symbol with a strict '<' scan (loss.py:236-246)
        using RT = typename RealOf<CT>::type;
        RT bv = (RT)-1;
        int bi = -1;
        for (int m = lane; m < g.M; m += 32) {
            const CT x = xmap_sec[m];
            const RT a = hypot(x.x, x.y);
            const bool an = a != a, bn = bv != bv;
            if ((an && (!bn || m > bi)) || (!an && !bn && (a > bv || (a == bv && m > bi)))) {
                bv = a;
                bi = m;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const RT ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const bool an = ov != ov, bn = bv != bv;
            if ((an && (!bn || oi > bi)) || (!an && !bn && (ov > bv || (ov == bv && oi > bi)))) {
                bv = ov;
                bi = oi;
            }
        }
        ant = bi;
        const CT x = xmap_sec[bi];
        double d = INFINITY;
        sym = -1;
        for (int k = 0; k < al.K; ++k) {
            const double ds = hypot((double)x.x - al.re[k], (double)x.y - al.im[k]);
            if (ds < d) {
                d = ds;
                sym = k;
            }
        }
    }
}

// Per-block accumulators, flushed to global memory once per kernel (keeps atomics off the frame loop).
struct BlockCounters {
    unsigned long long c[C_NUM_INT];
    double sq[4];
};

__device__ inline void counters_reset(BlockCounters* bc) {
    for (int i = threadIdx.x; i < C_NUM_INT; i += blockDim.x) bc->c[i] = 0ull;
    for (int i = threadIdx.x; i < 4; i += blockDim.x) bc->sq[i] = 0.0;
}
__device__ inline void counters_flush(const BlockCounters* bc, unsigned long long* out) {
    for (int i = threadIdx.x; i < C_NUM_INT; i += blockDim.x)
        if (bc->c[i]) atomicAdd(out + i, bc->c[i]);
    for (int i = threadIdx.x; i < 4; i += blockDim.x)
        if (bc->sq[i] != 0.0) atomicAdd(reinterpret_cast<double*>(out) + C_SQERR + i, bc->sq[i]);
}

// Loss epilogue of one frame (block-cooperative).  xmap / xmmse: the frame's estimates (shared or global),
// x_true etc. in global memory.  flags: shared int[4 + Lin] scratch.  Ends with a block barrier.
template <typename CT>
__device__ inline void block_loss(const Geom& g, const DevAlphabet& al, long long frame, const CT* xmap,
                                  const float2* xmmse, const LossIO& io, int iters, BlockCounters* bc, int* flags) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int mid = g.Lin / 2;
    // flags[0] = NaN seen, flags[1 + slot] = slot has a wrong entry
    for (int i = threadIdx.x; i < 1 + g.Lin; i += blockDim.x) flags[i] = 0;
    __syncthreads();
    const float2* xt = io.x_true + frame * (long long)g.N;
    double sq_all = 0.0, sq_f = 0.0, sq_m = 0.0, sq_l = 0.0;
    unsigned long long idx_err = 0, sym_err = 0, ibit = 0, sbit = 0;
    if (g.decision == 2) {
        // ---- 'random' mode (loss.py:252-280): per time slot the Na entries of largest |x| (np.abs(x).argsort()[-Na:],
        // NaN sorts last = largest), each decided to its nearest symbol (first minimum, strict '<', complex128 distance);
        // decided positions in ascending order pair with the true ones (loss.py:278, 165-172).  One warp per slot.
        for (int slot = warp; slot < g.Lin; slot += nwarps) {
            const int base = slot * g.Nt;
            unsigned taken = 0u;                                     // bit i: entry lane + 32 i already picked
            int my_pos = -1, my_k = -1;                              // lane a < Na holds the a-th pick
            for (int a = 0; a < g.Na; ++a) {
                float bv = -1.f;
                int bi = -1;
                for (int i = 0; lane + 32 * i < g.Nt; ++i) {
                    if (taken & (1u << i)) continue;
                    const int j = lane + 32 * i;
                    const CT x = xmap[base + j];
                    float v = hypotf((float)x.x, (float)x.y);
                    if (v != v) v = INFINITY;                        // NaN sorts last
                    if (v > bv || (v == bv && j > bi)) { bv = v; bi = j; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (ov > bv || (ov == bv && oi > bi)) { bv = ov; bi = oi; }
                }
                if (bi >= 0 && (bi & 31) == lane) taken |= 1u << (bi >> 5);
                if (lane == a) my_pos = bi;
            }
            if (lane < g.Na && my_pos >= 0) {
                const CT x = xmap[base + my_pos];
                double d = INFINITY;
                for (int k = 0; k < al.K; ++k) {
                    const double ds = hypot((double)x.x - al.re[k], (double)x.y - al.im[k]);
                    if (ds < d) { d = ds; my_k = k; }
                }
                if (my_k < 0) my_pos = -1;                           // no finite distance: the entry stays empty
            }
            // ascending order of the decided positions: rank by counting (empty picks go last)
            const int key = (lane < g.Na && my_pos >= 0) ? my_pos : 0x7fffffff;
            int rank = 0;
            for (int o = 0; o < g.Na; ++o) {
                const int other = __shfl_sync(0xffffffffu, key, o);
                rank += (other < key) || (other == key && o < lane);
            }
            if (lane < g.Na && my_pos >= 0) {
                // (the reference pairs the sorted decided list with the sorted true list over the WHOLE call; a slot with
                // fewer than Na decisions -- only possible with NaN estimates -- would misalign its arrays and numpy raises;
                // here such a slot is paired position by position within the slot)
                const long long ih = (g.frame_base + frame) * (long long)g.N + base + my_pos;
                const long long it = io.idx_true[(frame * g.Lin + slot) * g.Na + rank];
                const long long sh = al.gray[my_k], st = io.sym_true[(frame * g.Lin + slot) * g.Na + rank];
                idx_err += (ih != it);
                sym_err += (sh != st);
                const unsigned long long imask = g.index_bits_kept >= 64 ? ~0ull : ((1ull << g.index_bits_kept) - 1ull);
                ibit += __popcll((unsigned long long)(ih ^ it) & imask);
                sbit += __popcll((unsigned long long)(sh ^ st) & ((1ull << al.sbits) - 1ull));
            }
            // value compare of the decided slot against x, squared error of the MMSE estimate (all lanes run the shuffles)
            bool wrong = false, nan_seen = false;
            for (int j0 = 0; j0 < g.Nt; j0 += 32) {
                const int j = j0 + lane;
                float2 h = make_float2(0.f, 0.f);
                for (int o = 0; o < g.Na; ++o) {
                    const int pos = __shfl_sync(0xffffffffu, my_pos, o), kk = __shfl_sync(0xffffffffu, my_k, o);
                    if (pos == j && kk >= 0) h = make_float2((float)al.re[kk], (float)al.im[kk]);
                }
                if (j < g.Nt) {
                    const float2 t = xt[base + j];
                    wrong |= (h.x != t.x) || (h.y != t.y);
                    const float2 e = xmmse[base + j];
                    const float dr = e.x - t.x, di = e.y - t.y;
                    const double se = (double)dr * dr + (double)di * di;
                    sq_all += se;
                    if (slot == 0) sq_f += se;
                    if (slot == mid) sq_m += se;
                    if (slot == g.Lin - 1) sq_l += se;
                    const CT xm = xmap[base + j];
                    nan_seen |= (xm.x != xm.x) || (xm.y != xm.y);
                }
            }
            if (__any_sync(0xffffffffu, wrong) && lane == 0) flags[1 + slot] = 1;
            if (__any_sync(0xffffffffu, nan_seen) && lane == 0) flags[0] = 1;
        }
    } else
    for (int sec = warp; sec < g.L; sec += nwarps) {
        const int base = sec * g.M;
        const int slot = sec / g.Na;
        int ant, k;
        warp_decide(g, al, xmap + base, ant, k);
        const float2 shat = k >= 0 ? make_float2((float)al.re[k], (float)al.im[k]) : make_float2(0.f, 0.f);
        bool wrong = false, nan_seen = false;
        for (int m = lane; m < g.M; m += 32) {
            const float2 t = xt[base + m];
            const float2 h = (m == ant) ? shat : make_float2(0.f, 0.f);
            wrong |= (h.x != t.x) || (h.y != t.y);
            const float2 e = xmmse[base + m];
            const float dr = e.x - t.x, di = e.y - t.y;
            const double se = (double)dr * dr + (double)di * di;
            sq_all += se;
            if (slot == 0) sq_f += se;
            if (slot == mid) sq_m += se;
            if (slot == g.Lin - 1) sq_l += se;
            const CT xm = xmap[base + m];
            nan_seen |= (xm.x != xm.x) || (xm.y != xm.y);
        }
        if (__any_sync(0xffffffffu, wrong) && lane == 0) flags[1 + slot] = 1;
        if (__any_sync(0xffffffffu, nan_seen) && lane == 0) flags[0] = 1;
        if (lane == 0 && k >= 0) {
            const long long ih = (g.frame_base + frame) * (long long)g.N + base + ant;
            const long long it = io.idx_true[frame * g.L + sec];
            const long long sh = al.gray[k], st = io.sym_true[frame * g.L + sec];
            idx_err += (ih != it);
            sym_err += (sh != st);
            const unsigned long long imask = g.index_bits_kept >= 64 ? ~0ull : ((1ull << g.index_bits_kept) - 1ull);
            ibit += __popcll((unsigned long long)(ih ^ it) & imask);
            sbit += __popcll((unsigned long long)(sh ^ st) & ((1ull << al.sbits) - 1ull));
        }
    }
    sq_all = warp_sum(sq_all);
    sq_f = warp_sum(sq_f);
    sq_m = warp_sum(sq_m);
    sq_l = warp_sum(sq_l);
    if (g.decision == 2) {          // random mode books its label counters on lanes 0..Na-1
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            idx_err += __shfl_xor_sync(0xffffffffu, idx_err, o);
            sym_err += __shfl_xor_sync(0xffffffffu, sym_err, o);
            ibit += __shfl_xor_sync(0xffffffffu, ibit, o);
            sbit += __shfl_xor_sync(0xffffffffu, sbit, o);
        }
    }
    if (lane == 0) {
        atomicAdd(&bc->sq[0], sq_all);
        atomicAdd(&bc->sq[1], sq_f);
        atomicAdd(&bc->sq[2], sq_m);
        atomicAdd(&bc->sq[3], sq_l);
        if (idx_err) atomicAdd(&bc->c[C_INDEX_ERR], idx_err);
        if (sym_err) atomicAdd(&bc->c[C_SYMBOL_ERR], sym_err);
        if (ibit) atomicAdd(&bc->c[C_INDEX_BIT], ibit);
        if (sbit) atomicAdd(&bc->c[C_SYMBOL_BIT], sbit);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int bad_slots = 0;
        for (int s = 0; s < g.Lin; ++s) bad_slots += flags[1 + s];
        bc->c[C_FRAMES] += 1;
        bc->c[C_FRAME_ERR] += bad_slots > 0;
        bc->c[C_SLOT_ERR] += bad_slots;
        bc->c[C_SLOT_FIRST] += flags[1];
        bc->c[C_SLOT_MID] += flags[1 + mid];
        bc->c[C_SLOT_LAST] += flags[g.Lin];
        bc->c[C_ITERS] += iters;
        bc->c[C_NAN_FRAMES] += flags[0];
    }
    __syncthreads();
}

}  // namespace ampsm
