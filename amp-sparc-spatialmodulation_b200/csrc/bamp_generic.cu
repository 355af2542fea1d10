// Generic BAMP kernel: one CTA per frame (persistent over frames), any n x N whose vectors fit shared memory.
// The frame's channel matrix is staged into shared memory once with a 1-D bulk TMA copy (cp.async.bulk +
// mbarrier) and all iterations run in-kernel; matrices that do not fit are read through L2 instead.
// Follows bamp.py:12-25 (state), 59-64 (iteration), 66-77 (denoiser), 116-143 (loop, exit, Loss).
//
// Structured operator (OPK > 0, ampsm_bamp_detect_taps): for ISI channels (Lh > 1, Lin > 1) the reference's matrix is
// block-Toeplitz -- block (i, j) of H is the Nr x Nt tap matrix g[i - j] (channel.py:53-72, 85-91).  Only the taps
// [Lh][Nr][Nt] live in shared memory (rows padded by one element: conflict-free for both passes) and H, H^H, |H|^2,
// |H|^2^T are applied as block convolutions: a thread owns TI consecutive time slots of one antenna, so every tap it
// loads is used TI times and there are no cross-lane reductions.  Slots outside the frame point at a zeroed slot.
#include <cstdlib>

#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

struct BampPlan {
    size_t H, y, z, g, u, w, xh, xmap, var, var_new, cov, scr, red, flags, bc, mbar, total;
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

__host__ __device__ inline size_t taps_ld(const Geom& g) { return (size_t)g.Nt + 1; }

__host__ __device__ inline BampPlan bamp_plan(const Geom& g, bool stage, bool exp64, int Lh = 0) {
    BampPlan p;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align16(o + bytes);
        return at;
    };
    // operator: the dense matrix, or Lh tap matrices with padded rows; vectors read by the block convolutions carry one
    // extra all-zero time slot (Nr resp. Nt entries)
    p.H = take(!stage ? 0 : (Lh > 0 ? (size_t)Lh * g.Nr * taps_ld(g) * 8 : (size_t)g.n * g.N * 8));
    const size_t nz = Lh > 0 ? (size_t)g.Nr : 0, Nz = Lh > 0 ? (size_t)g.Nt : 0;
    p.y = take((size_t)g.n * 8);
    p.z = take((size_t)g.n * 8);
    p.g = take(((size_t)g.n + nz) * 8);
    p.u = take((size_t)g.n * 4);
    p.w = take(((size_t)g.n + nz) * 4);
    p.xh = take(((size_t)g.N + Nz) * 8);
    p.xmap = take((size_t)g.N * 8);
    p.var = take(((size_t)g.N + Nz) * 4);
    p.var_new = take((size_t)g.N * 4);
    p.cov = take((size_t)g.N * 4);
    p.scr = take(denoise_scratch_elems(g) * (exp64 ? 8 : 4));
    p.red = take(32 * 3 * 8);
    p.flags = take((size_t)(1 + g.Lin) * 4);
    p.bc = take(sizeof(BlockCounters));
    p.mbar = take(8);
    p.total = o;
    return p;
}

// block-wide sum of three doubles (trajectory means); result valid in every thread
__device__ inline void block_sum3(double& a, double& b, double& c, double* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
        red[warp] = a;
        red[32 + warp] = b;
        red[64 + warp] = c;
    }
    __syncthreads();
    a = b = c = 0.0;
    for (int w = 0; w < nw; ++w) {
        a += red[w];
        b += red[32 + w];
        c += red[64 + w];
    }
    __syncthreads();
}

constexpr int kWinLh = 4;   // channel lengths up to this use the sliding-window loops of the structured operator

// OPK: 0 dense matrix; otherwise the structured operator with OPK time slots per thread
template <bool EXP64, int OPK>
__global__ void __launch_bounds__(512) bamp_generic_kernel(const __grid_constant__ BampArgs a) {
    using E = typename ExpT<EXP64>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    const bool stage = a.stage_H != 0;
    const BampPlan P = bamp_plan(g, stage, EXP64, OPK ? a.Lh : 0);
    float2* Hs = reinterpret_cast<float2*>(smem + P.H);
    float2* y_s = reinterpret_cast<float2*>(smem + P.y);
    float2* z_s = reinterpret_cast<float2*>(smem + P.z);
    float2* g_s = reinterpret_cast<float2*>(smem + P.g);
    float* u_s = reinterpret_cast<float*>(smem + P.u);
    float* w_s = reinterpret_cast<float*>(smem + P.w);
    float2* xh_s = reinterpret_cast<float2*>(smem + P.xh);
    float2* xmap_s = reinterpret_cast<float2*>(smem + P.xmap);
    float* var_s = reinterpret_cast<float*>(smem + P.var);
    float* varn_s = reinterpret_cast<float*>(smem + P.var_new);
    float* cov_s = reinterpret_cast<float*>(smem + P.cov);
    E* scr = reinterpret_cast<E*>(smem + P.scr);
    double* red = reinterpret_cast<double*>(smem + P.red);
    int* flags = reinterpret_cast<int*>(smem + P.flags);
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem + P.bc);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + P.mbar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n = g.n, N = g.N;
    const uint32_t Hbytes = (uint32_t)((size_t)n * N * 8);
    const bool shared_H = a.H_stride == 0;
    const int ldt = OPK ? (stage ? (int)taps_ld(g) : g.Nt) : 0;          // row stride of a tap matrix

    counters_reset(bc);
    if (OPK) {                                                              // the zero slots (never written again)
        for (int i = tid; i < g.Nr; i += blockDim.x) { g_s[n + i] = make_float2(0.f, 0.f); w_s[n + i] = 0.f; }
        for (int j = tid; j < g.Nt; j += blockDim.x) { xh_s[N + j] = make_float2(0.f, 0.f); var_s[N + j] = 0.f; }
    }
    if (!OPK && stage && tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    bool H_loaded = false;

    for (long long f = blockIdx.x; f < a.frames; f += gridDim.x) {
        const float2* Hg = a.H + f * a.H_stride;
        const float2* Hm = stage ? Hs : Hg;
        if (stage && !(shared_H && H_loaded)) {
            // all threads finished reading the previous frame's matrix (barrier at the end of the last frame)
            if (OPK) {
                const int rows = a.Lh * g.Nr;
                for (int e = tid; e < rows * g.Nt; e += blockDim.x) Hs[(size_t)(e / g.Nt) * ldt + e % g.Nt] = Hg[e];
                H_loaded = true;
            } else if (tid == 0) {
                mbar_expect_tx(mbar, Hbytes);
                tma_load_1d(Hs, Hg, Hbytes, mbar);
            }
        }
        const float sigma2 = a.sigma2_pf ? a.sigma2_pf[f] : a.sigma2;
        for (int i = tid; i < n; i += blockDim.x) {
            const float2 yv = a.y[f * n + i];
            y_s[i] = yv;
            z_s[i] = yv;          // z = y            (bamp.py:23)
            u_s[i] = sigma2;      // u = v + sigma2, v = 0 (bamp.py:25)
        }
        for (int j = tid; j < N; j += blockDim.x) {
            xh_s[j] = make_float2(0.f, 0.f);   // bamp.py:20
            var_s[j] = 1.0f;                   // bamp.py:21
        }
        if (!OPK && stage && !(shared_H && H_loaded)) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            H_loaded = true;
        }
        __syncthreads();

        int t_done = 0;
        for (int t = 0; t < g.max_iters; ++t) {
            // ---- row pass: v = |H|^2 var, Hx = H xhat; then z, u and the column pass operands (bamp.py:59-61)
            auto row_update = [&](int i, float av, float ar, float ai) {
                const float2 yv = y_s[i], zo = z_s[i];
                const float2 resid = make_float2(yv.x - zo.x, yv.y - zo.y);
                const float2 corr = cdiv_real(make_float2(av * resid.x, av * resid.y), u_s[i]);   // old u
                const float2 zn = make_float2(ar - corr.x, ai - corr.y);
                const float un = av + sigma2;
                z_s[i] = zn;
                u_s[i] = un;
                g_s[i] = cdiv_real(make_float2(yv.x - zn.x, yv.y - zn.y), un);
                w_s[i] = __frcp_rn(un);
            };
            if constexpr (OPK == 0) {
                for (int i = warp; i < n; i += nwarps) {
                    const float2* Hrow = Hm + (size_t)i * N;
                    float av = 0.f, ar = 0.f, ai = 0.f;
                    for (int j = lane; j < N; j += 32) {
                        const float2 h = Hrow[j];
                        const float2 x = xh_s[j];
                        av = fmaf(fmaf(h.x, h.x, h.y * h.y), var_s[j], av);
                        ar = fmaf(h.x, x.x, fmaf(-h.y, x.y, ar));
                        ai = fmaf(h.x, x.y, fmaf(h.y, x.x, ai));
                    }
                    av = warp_sum(av);
                    ar = warp_sum(ar);
                    ai = warp_sum(ai);
                    if (lane == 0) row_update(i, av, ar, ai);
                }
            } else {
                // out(i, r) = sum_l sum_t g_l[r][t] x(i - l, t): thread = (tile of OPK output slots, receive antenna r)
                const int tiles = (g.Lout + OPK - 1) / OPK;
                for (int w = tid; w < tiles * g.Nr; w += blockDim.x) {
                    const int r = w % g.Nr, i0 = (w / g.Nr) * OPK;
                    float av[OPK], ar[OPK], ai[OPK];
#pragma unroll
                    for (int k = 0; k < OPK; ++k) av[k] = ar[k] = ai[k] = 0.f;
                    if (a.Lh <= kWinLh) {
                        // sliding window: the OPK outputs and Lh taps touch only OPK + Lh - 1 input slots; each is loaded
                        // once per column (window entry q = k - l + kWinLh - 1 is input slot i0 + q - (kWinLh - 1))
                        constexpr int W = OPK + kWinLh - 1;
                        int jb[W];
#pragma unroll
                        for (int q = 0; q < W; ++q) {
                            int j = i0 + q - (kWinLh - 1);
                            if (a.cyclic && j < 0) j += g.Lin;
                            jb[q] = (j >= 0 && j < g.Lin ? j : g.Lin) * g.Nt;
                        }
                        const int qmin = kWinLh - a.Lh;
                        const float2* trow = Hm + (size_t)r * ldt;
                        const size_t tap_stride = (size_t)g.Nr * ldt;
                        for (int c = 0; c < g.Nt; ++c) {
                            float2 h[kWinLh], xw[W];
                            float p[kWinLh], vw[W];
#pragma unroll
                            for (int l = 0; l < kWinLh; ++l)
                                if (l < a.Lh) {
                                    h[l] = trow[l * tap_stride + c];
                                    p[l] = fmaf(h[l].x, h[l].x, h[l].y * h[l].y);
                                }
#pragma unroll
                            for (int q = 0; q < W; ++q)
                                if (q >= qmin) {
                                    xw[q] = xh_s[jb[q] + c];
                                    vw[q] = var_s[jb[q] + c];
                                }
#pragma unroll
                            for (int l = 0; l < kWinLh; ++l)
                                if (l < a.Lh) {
#pragma unroll
                                    for (int k = 0; k < OPK; ++k) {
                                        const int q = k - l + kWinLh - 1;
                                        av[k] = fmaf(p[l], vw[q], av[k]);
                                        ar[k] = fmaf(h[l].x, xw[q].x, fmaf(-h[l].y, xw[q].y, ar[k]));
                                        ai[k] = fmaf(h[l].x, xw[q].y, fmaf(h[l].y, xw[q].x, ai[k]));
                                    }
                                }
                        }
                    } else {
                        for (int l = 0; l < a.Lh; ++l) {
                            int js[OPK];                                   // first entry of the input slot (zero slot if none)
#pragma unroll
                            for (int k = 0; k < OPK; ++k) {
                                int j = i0 + k - l;
                                if (a.cyclic && j < 0) j += g.Lin;
                                js[k] = (j >= 0 && j < g.Lin ? j : g.Lin) * g.Nt;
                            }
                            const float2* trow = Hm + (size_t)(l * g.Nr + r) * ldt;
                            for (int c = 0; c < g.Nt; ++c) {
                                const float2 h = trow[c];
                                const float p = fmaf(h.x, h.x, h.y * h.y);
#pragma unroll
                                for (int k = 0; k < OPK; ++k) {
                                    const float2 x = xh_s[js[k] + c];
                                    av[k] = fmaf(p, var_s[js[k] + c], av[k]);
                                    ar[k] = fmaf(h.x, x.x, fmaf(-h.y, x.y, ar[k]));
                                    ai[k] = fmaf(h.x, x.y, fmaf(h.y, x.x, ai[k]));
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < OPK; ++k)
                        if (i0 + k < g.Lout) row_update((i0 + k) * g.Nr + r, av[k], ar[k], ai[k]);
                }
            }
            __syncthreads();
            // ---- column pass: cov = 1/(|H|^2^T (1/u)), xmap = xhat + cov * H^H((y-z)/u) (bamp.py:62-63)
            auto col_update = [&](int j, float ac, float ar, float ai) {
                const float cov = __frcp_rn(ac);
                const float2 x = xh_s[j];
                xmap_s[j] = make_float2(fmaf(cov, ar, x.x), fmaf(cov, ai, x.y));
                cov_s[j] = cov;
            };
            if constexpr (OPK == 0) {
                for (int j = tid; j < N; j += blockDim.x) {
                    float ac = 0.f, ar = 0.f, ai = 0.f;
                    for (int i = 0; i < n; ++i) {
                        const float2 h = Hm[(size_t)i * N + j];
                        const float2 gv = g_s[i];
                        ac = fmaf(fmaf(h.x, h.x, h.y * h.y), w_s[i], ac);
                        ar = fmaf(h.x, gv.x, fmaf(h.y, gv.y, ar));
                        ai = fmaf(h.x, gv.y, fmaf(-h.y, gv.x, ai));
                    }
                    col_update(j, ac, ar, ai);
                }
            } else {
                // out(j, t) = sum_l sum_r conj(g_l[r][t]) q(j + l, r): thread = (tile of OPK input slots, transmit antenna t)
                const int tiles = (g.Lin + OPK - 1) / OPK;
                for (int w = tid; w < tiles * g.Nt; w += blockDim.x) {
                    const int tx = w % g.Nt, j0 = (w / g.Nt) * OPK;
                    float ac[OPK], ar[OPK], ai[OPK];
#pragma unroll
                    for (int k = 0; k < OPK; ++k) ac[k] = ar[k] = ai[k] = 0.f;
                    if (a.Lh <= kWinLh) {
                        // sliding window over the output slots j0 .. j0 + OPK + Lh - 2 (window entry q = k + l)
                        constexpr int W = OPK + kWinLh - 1;
                        int ib[W];
#pragma unroll
                        for (int q = 0; q < W; ++q) {
                            int i = j0 + q;
                            if (a.cyclic && i >= g.Lin) i -= g.Lin;
                            ib[q] = (i < g.Lout ? i : g.Lout) * g.Nr;
                        }
                        const int qend = OPK + a.Lh - 1;
                        const float2* tcol = Hm + tx;
                        const size_t tap_stride = (size_t)g.Nr * ldt;
                        for (int r = 0; r < g.Nr; ++r) {
                            float2 h[kWinLh], gw[W];
                            float p[kWinLh], ww[W];
#pragma unroll
                            for (int l = 0; l < kWinLh; ++l)
                                if (l < a.Lh) {
                                    h[l] = tcol[l * tap_stride + (size_t)r * ldt];
                                    p[l] = fmaf(h[l].x, h[l].x, h[l].y * h[l].y);
                                }
#pragma unroll
                            for (int q = 0; q < W; ++q)
                                if (q < qend) {
                                    gw[q] = g_s[ib[q] + r];
                                    ww[q] = w_s[ib[q] + r];
                                }
#pragma unroll
                            for (int l = 0; l < kWinLh; ++l)
                                if (l < a.Lh) {
#pragma unroll
                                    for (int k = 0; k < OPK; ++k) {
                                        const int q = k + l;
                                        ac[k] = fmaf(p[l], ww[q], ac[k]);
                                        ar[k] = fmaf(h[l].x, gw[q].x, fmaf(h[l].y, gw[q].y, ar[k]));
                                        ai[k] = fmaf(h[l].x, gw[q].y, fmaf(-h[l].y, gw[q].x, ai[k]));
                                    }
                                }
                        }
                    } else {
                        for (int l = 0; l < a.Lh; ++l) {
                            int is[OPK];
#pragma unroll
                            for (int k = 0; k < OPK; ++k) {
                                int i = j0 + k + l;
                                if (a.cyclic && i >= g.Lin) i -= g.Lin;
                                is[k] = (i < g.Lout ? i : g.Lout) * g.Nr;
                            }
                            const float2* tcol = Hm + (size_t)l * g.Nr * ldt + tx;
                            for (int r = 0; r < g.Nr; ++r) {
                                const float2 h = tcol[(size_t)r * ldt];
                                const float p = fmaf(h.x, h.x, h.y * h.y);
#pragma unroll
                                for (int k = 0; k < OPK; ++k) {
                                    const float2 gv = g_s[is[k] + r];
                                    ac[k] = fmaf(p, w_s[is[k] + r], ac[k]);
                                    ar[k] = fmaf(h.x, gv.x, fmaf(h.y, gv.y, ar[k]));
                                    ai[k] = fmaf(h.x, gv.y, fmaf(-h.y, gv.x, ai[k]));
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < OPK; ++k)
                        if (j0 + k < g.Lin) col_update((j0 + k) * g.Nt + tx, ac[k], ar[k], ai[k]);
                }
            }
            __syncthreads();
            // ---- denoiser (bamp.py:66-77): tau = cov/2
            if (g.decision == 2) {                                   // generator_mode 'random' (bamp.py:46): i.i.d. prior
                block_denoise_iid(g, al, xmap_s, cov_s, xh_s, varn_s);
            } else {
                double gshift = 0.0;
                if (EXP64 && g.shift_mode == 1) gshift = block_absmax_exponent(g, al, xmap_s, cov_s, 0.f, true, red);
                block_denoise<EXP64>(g, al, xmap_s, cov_s, 0.f, true, gshift, xh_s, varn_s, scr, 1, denoise_scratch_per_warp(g));
            }
            __syncthreads();
            // ---- exit test on var (bamp.py:140) and state update
            bool close = true;
            for (int j = tid; j < N; j += blockDim.x) {
                const float vn = varn_s[j], vo = var_s[j];
                close &= fabsf(vn - vo) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, vo)));
            }
            const int all_close = __syncthreads_and(close ? 1 : 0);
            if (a.traj) {
                double s_tau = 0.0, s_var = 0.0, s_mse = 0.0;
                for (int j = tid; j < N; j += blockDim.x) {
                    s_tau += cov_s[j];
                    s_var += varn_s[j];
                    if (a.io.x_true) {
                        const float2 xt = a.io.x_true[f * N + j], xe = xh_s[j];
                        const double dr = (double)xe.x - xt.x, di = (double)xe.y - xt.y;
                        s_mse += dr * dr + di * di;
                    }
                }
                block_sum3(s_tau, s_var, s_mse, red);
                if (tid == 0) {
                    float* tr = a.traj + (f * g.max_iters + t) * 3;
                    tr[0] = (float)(s_tau / N);
                    tr[1] = (float)(s_var / N);
                    tr[2] = (float)(s_mse / N);
                }
            }
            for (int j = tid; j < N; j += blockDim.x) var_s[j] = varn_s[j];
            __syncthreads();
            t_done = t + 1;
            if (g.early_exit && all_close) break;
        }
        // ---- outputs and Loss (bamp.py:142)
        for (int j = tid; j < N; j += blockDim.x) {
            if (a.xmap) a.xmap[f * N + j] = xmap_s[j];
            if (a.xmmse) a.xmmse[f * N + j] = xh_s[j];
            if (a.var) a.var[f * N + j] = var_s[j];
        }
        if (a.traj) {   // repeat the last value for the iterations that did not run
            for (int t = t_done + tid; t < g.max_iters; t += blockDim.x) {
                const float* last = a.traj + (f * g.max_iters + t_done - 1) * 3;
                float* tr = a.traj + (f * g.max_iters + t) * 3;
                tr[0] = last[0];
                tr[1] = last[1];
                tr[2] = last[2];
            }
        }
        if (tid == 0 && a.iters) a.iters[f] = t_done;
        if (a.io.x_true) {
            block_loss(g, al, f, xmap_s, xh_s, a.io, t_done, bc, flags);
        } else if (tid == 0) {
            bc->c[C_FRAMES] += 1;
            bc->c[C_ITERS] += t_done;
        }
        __syncthreads();
    }
    __syncthreads();
    if (a.io.counters) counters_flush(bc, a.io.counters);
}

int launch_bamp_generic(const BampArgs& args, bool exp64, cudaStream_t stream) {
    int dev = 0, sms = 0, smem_max = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    BampArgs a = args;
    const Geom& g = a.g;
    const int Lh = a.taps ? a.Lh : 0;
    if (Lh) { a.H = a.taps; a.H_stride = a.taps_stride; }
    // dense: bulk TMA needs 16-byte aligned, 16-byte multiple transfers; taps are staged by plain loads
    const bool stage_ok = Lh ? true
                             : ((size_t)g.n * g.N * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.H) % 16) == 0 &&
                                   ((size_t)a.H_stride * 8) % 16 == 0 && (size_t)g.n * g.N * 8 < (1u << 20);
    BampPlan plan = bamp_plan(g, stage_ok, exp64, Lh);
    a.stage_H = stage_ok && plan.total <= (size_t)smem_max;
    if (!a.stage_H) plan = bamp_plan(g, false, exp64, Lh);
    if (plan.total > (size_t)smem_max) {
        set_error("BAMP generic kernel: per-frame vectors need %zu B of shared memory (> %d B)", plan.total, smem_max);
        return AMPSM_ENOFIT;
    }
    int threads = g.N >= 128 ? 256 : (g.N >= 64 ? 128 : ((long long)g.n * g.N <= 64 ? 32 : 64));
    // structured operator: frames this large leave one CTA per SM (shared memory), so give it 16 warps; then the most time
    // slots per thread (tap reuse) that still give 60 % of the threads a work item in the row pass
    if (Lh && g.N >= 2048) threads = 512;
    if (Lh && getenv("AMPSM_TAPS_THREADS")) threads = atoi(getenv("AMPSM_TAPS_THREADS"));      // tuning switches
    int opk = 0;
    if (Lh) {
        opk = 1;
        for (int k : {4, 2})
            if (((g.Lout + k - 1) / k) * g.Nr * 5 >= threads * 3) { opk = k; break; }
        if (getenv("AMPSM_TAPS_SLOTS")) opk = atoi(getenv("AMPSM_TAPS_SLOTS"));
        if (opk != 1 && opk != 2 && opk != 4) opk = 1;
    }
    void (*kern)(const BampArgs);
    switch (opk) {
        case 0: kern = exp64 ? bamp_generic_kernel<true, 0> : bamp_generic_kernel<false, 0>; break;
        case 1: kern = exp64 ? bamp_generic_kernel<true, 1> : bamp_generic_kernel<false, 1>; break;
        case 2: kern = exp64 ? bamp_generic_kernel<true, 2> : bamp_generic_kernel<false, 2>; break;
        default: kern = exp64 ? bamp_generic_kernel<true, 4> : bamp_generic_kernel<false, 4>; break;
    }
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total),
                           "cudaFuncSetAttribute(bamp_generic)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, plan.total);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > a.frames) grid = a.frames;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, plan.total, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "bamp_generic_kernel launch");
}

}  // namespace ampsm
