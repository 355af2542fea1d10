// vamp2.py of the reference -- "direct implementation of Rangan (with damping)", the second VAMP that no driver of the reference
// imports -- as one CTA per frame, Vh staged in shared memory by one bulk TMA copy per frame.
// Follows vamp2.py:12-26 (Tracker: y~ = U^H y / s, gamma = 1, r = 0, var = 1, eta = N / R), 52-76 (one layer: denoiser with the
// scalar precision-like `gamma` as tau, damped posterior mean, alpha = mean(var) gamma, r~ and gamma~ with the 1e-11 / 1e11
// clips, d = s^2 / (s^2 + sigma^2 gamma~), gamma <- damped gamma~ mean(d) / (eta - mean(d)), r = r~ + eta V (d / mean(d) (y~ - Vh r~))),
// 78-87 (denoiser: vamp.py's soft-max, variance E|s|^2 - |E s|^2) and 117-127 (loop, allclose exit on var, Loss on (r, xmmse)).
// complex64 only (the class is never fed anything else); lines 73-74 of the layer are dead code and have no counterpart here.
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

namespace {

struct Vamp2Plan {
    size_t Vh, yt, e, s2, rt, r, xh, xm, var, var_new, scr, red, flags, bc, mbar, total;
};
__host__ __device__ inline Vamp2Plan vamp2_plan(const Geom& g, bool stage, bool exp64) {
    Vamp2Plan p;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = (o + bytes + 15) & ~size_t(15);
        return at;
    };
    p.Vh = take(stage ? (size_t)g.R * g.N * 8 : 0);
    p.yt = take((size_t)g.R * 8);
    p.e = take((size_t)g.R * 8);
    p.s2 = take((size_t)g.R * 4);
    p.rt = take((size_t)g.N * 8);
    p.r = take((size_t)g.N * 8);
    p.xh = take((size_t)g.N * 8);
    p.xm = take((size_t)g.N * 8);
    p.var = take((size_t)g.N * 4);
    p.var_new = take((size_t)g.N * 4);
    p.scr = take((size_t)g.N * 3 * (exp64 ? 8 : 4));
    p.red = take(32 * 3 * 8);
    p.flags = take((size_t)(1 + g.Lin) * 4);
    p.bc = take(sizeof(BlockCounters));
    p.mbar = take(8);
    p.total = o;
    return p;
}

// torch.max / torch.min propagate NaN (vamp2.py:67-68)
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
    if (v != v) return v;
    return fminf(fmaxf(v, lo), hi);
}

__device__ inline double block_sum2(double a, double* red) {
    a = warp_sum(a);
    const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = a;
    __syncthreads();
    double r = 0.0;
    for (int w = 0; w < nw; ++w) r += red[w];
    __syncthreads();
    return r;
}

template <bool EXP64>
__global__ void __launch_bounds__(256) vamp2_kernel(const __grid_constant__ VampArgs a) {
    using E = typename ExpT<EXP64>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    const bool stage = a.stage_Vh != 0;
    const Vamp2Plan P = vamp2_plan(g, stage, EXP64);
    float2* Vs = reinterpret_cast<float2*>(smem + P.Vh);
    float2* yt_s = reinterpret_cast<float2*>(smem + P.yt);
    float2* e_s = reinterpret_cast<float2*>(smem + P.e);
    float* s2_s = reinterpret_cast<float*>(smem + P.s2);
    float2* rt_s = reinterpret_cast<float2*>(smem + P.rt);
    float2* r_s = reinterpret_cast<float2*>(smem + P.r);
    float2* xh_s = reinterpret_cast<float2*>(smem + P.xh);
    float2* xm_s = reinterpret_cast<float2*>(smem + P.xm);
    float* var_s = reinterpret_cast<float*>(smem + P.var);
    float* varn_s = reinterpret_cast<float*>(smem + P.var_new);
    E* scr = reinterpret_cast<E*>(smem + P.scr);
    double* red = reinterpret_cast<double*>(smem + P.red);
    int* flags = reinterpret_cast<int*>(smem + P.flags);
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem + P.bc);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + P.mbar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n = g.n, N = g.N, R = g.R;
    const uint32_t Vbytes = (uint32_t)((size_t)R * N * 8);
    const bool shared_V = a.Vh_stride == 0;
    const float eta = (float)((double)N / (double)R);                     // vamp2.py:26 (python float, float32 in every product)
    const float rho = a.damping, one_m_rho = (float)(1.0 - (double)a.damping);
    const float var_min = 1.0e-11f, var_max = 1.0e11f;                    // vamp2.py:48-49

    counters_reset(bc);
    if (stage && tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    bool V_loaded = false;

    for (long long f = blockIdx.x; f < a.frames; f += gridDim.x) {
        const float2* Vg = reinterpret_cast<const float2*>(a.Vh) + f * a.Vh_stride;
        const float2* Vm = stage ? Vs : Vg;
        const float2* Ug = reinterpret_cast<const float2*>(a.U) + f * a.U_stride;
        const float* sg = reinterpret_cast<const float*>(a.s) + f * a.s_stride;
        const float2* yg = reinterpret_cast<const float2*>(a.y) + f * n;
        if (stage && !(shared_V && V_loaded) && tid == 0) {
            mbar_expect_tx(mbar, Vbytes);
            tma_load_1d(Vs, Vg, Vbytes, mbar);
        }
        const float nv = (float)(a.sigma2_pf ? (double)a.sigma2_pf[f] : a.sigma2_d);
        // y~ = (U^H y) / s (vamp2.py:22), thread per singular value
        for (int k = tid; k < R; k += blockDim.x) {
            float ar = 0.f, ai = 0.f;
            for (int i = 0; i < n; ++i) {
                const float2 u = Ug[(size_t)i * R + k];
                const float2 yv = yg[i];
                ar += u.x * yv.x + u.y * yv.y;                      // conj(u) y
                ai += u.x * yv.y - u.y * yv.x;
            }
            const float sk = sg[k];
            yt_s[k] = make_float2(__fdiv_rn(ar, sk), __fdiv_rn(ai, sk));
            s2_s[k] = sk * sk;                                      // vamp2.py:17
        }
        for (int j = tid; j < N; j += blockDim.x) {
            r_s[j] = make_float2(0.f, 0.f);                         // vamp2.py:23-25
            var_s[j] = 1.0f;
            xh_s[j] = make_float2(0.f, 0.f);
        }
        float gamma = 1.0f;                                         // vamp2.py:21
        if (stage && !(shared_V && V_loaded)) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            V_loaded = true;
        }
        __syncthreads();

        int t_done = 0;
        for (int t = 0; t < g.max_iters; ++t) {
            // ---- (xmmse, var) = denoiser(r, gamma); xmmse <- rho xmmse + (1 - rho) xmmse_old  (vamp2.py:61-62)
            double gshift = 0.0;
            if (EXP64 && g.shift_mode == 1) gshift = block_absmax_exponent<float2>(g, al, r_s, nullptr, gamma, false, red);
            block_denoise<EXP64, float2>(g, al, r_s, nullptr, gamma, false, gshift, xm_s, varn_s, scr, 1, false, true);
            __syncthreads();
            double vsum = 0.0;
            bool close = true;
            for (int j = tid; j < N; j += blockDim.x) {
                const float vn = varn_s[j], vo = var_s[j];
                vsum += vn;
                close &= fabsf(vn - vo) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, vo)));
                const float2 xm = xm_s[j], xo = xh_s[j];
                xh_s[j] = make_float2(__fadd_rn(__fmul_rn(rho, xm.x), __fmul_rn(one_m_rho, xo.x)),
                                      __fadd_rn(__fmul_rn(rho, xm.y), __fmul_rn(one_m_rho, xo.y)));
            }
            const int all_close = __syncthreads_and(close ? 1 : 0);
            const double vtot = block_sum2(vsum, red);
            const float alpha = __fmul_rn((float)(vtot / N), gamma);            // vamp2.py:63
            const float one_m_alpha = 1.0f - alpha;
            // ---- r~ = (xmmse - alpha r) / (1 - alpha); gamma~ = gamma (1 - alpha) / alpha, clipped  (vamp2.py:65-68)
            for (int j = tid; j < N; j += blockDim.x) {
                const float2 xe = xh_s[j], rv = r_s[j];
                rt_s[j] = make_float2(__fdiv_rn(xe.x - __fmul_rn(alpha, rv.x), one_m_alpha), __fdiv_rn(xe.y - __fmul_rn(alpha, rv.y), one_m_alpha));
                var_s[j] = varn_s[j];
            }
            const float g_tilde = clamp_nan(__fdiv_rn(__fmul_rn(gamma, one_m_alpha), alpha), var_min, var_max);
            const float ng = __fmul_rn(nv, g_tilde);
            __syncthreads();
            // ---- d = s^2 / (s^2 + sigma^2 gamma~), residual y~ - Vh r~ (vamp2.py:70, 76), warp per row
            double dsum = 0.0;
            for (int k = warp; k < R; k += nwarps) {
                const float2* Vrow = Vm + (size_t)k * N;
                float ar = 0.f, ai = 0.f;
                for (int j = lane; j < N; j += 32) {
                    const float2 v = Vrow[j], x = rt_s[j];
                    ar += v.x * x.x - v.y * x.y;
                    ai += v.x * x.y + v.y * x.x;
                }
                ar = warp_sum(ar);
                ai = warp_sum(ai);
                if (lane == 0) {
                    const float d = __fdiv_rn(s2_s[k], s2_s[k] + ng);
                    const float2 yt = yt_s[k];
                    e_s[k] = make_float2(yt.x - ar, yt.y - ai);     // scaled by d / mean(d) below
                    reinterpret_cast<float*>(scr)[k] = d;           // the denoiser scratch is free here (3 N >= R values)
                    dsum += (double)d;
                }
            }
            const double dtot = block_sum2(dsum, red);
            const float dm = (float)(dtot / R);
            const float g_new = __fdiv_rn(__fmul_rn(g_tilde, dm), eta - dm);    // vamp2.py:71
            gamma = __fadd_rn(__fmul_rn(rho, g_new), __fmul_rn(one_m_rho, gamma));   // vamp2.py:72
            for (int k = tid; k < R; k += blockDim.x) {
                const float w = __fdiv_rn(reinterpret_cast<float*>(scr)[k], dm);
                e_s[k] = make_float2(w * e_s[k].x, w * e_s[k].y);
            }
            __syncthreads();
            // ---- r = r~ + (eta V) (d / mean(d) (y~ - Vh r~))  (vamp2.py:76: `T.eta * T.V @ ...` scales the matrix first)
            double mse = 0.0;
            for (int j = tid; j < N; j += blockDim.x) {
                float ar = 0.f, ai = 0.f;
                for (int k = 0; k < R; ++k) {
                    const float2 v = Vm[(size_t)k * N + j], ev = e_s[k];
                    const float vx = __fmul_rn(eta, v.x), vy = -__fmul_rn(eta, v.y);      // eta conj(v)
                    ar += vx * ev.x - vy * ev.y;
                    ai += vx * ev.y + vy * ev.x;
                }
                const float2 rt = rt_s[j];
                r_s[j] = make_float2(rt.x + ar, rt.y + ai);
                if (a.traj && a.io.x_true) {
                    const float2 xt = a.io.x_true[f * N + j], xe = xh_s[j];
                    const double dr = (double)xe.x - xt.x, di = (double)xe.y - xt.y;
                    mse += dr * dr + di * di;
                }
            }
            if (a.traj) {
                mse = block_sum2(mse, red);
                if (tid == 0) {
                    float* tr = a.traj + (f * g.max_iters + t) * 3;
                    tr[0] = gamma;
                    tr[1] = (float)(vtot / N);
                    tr[2] = (float)(mse / N);
                }
            }
            __syncthreads();
            t_done = t + 1;
            if (g.early_exit && all_close) break;
        }
        for (int j = tid; j < N; j += blockDim.x) {
            if (a.xmap) reinterpret_cast<float2*>(a.xmap)[f * N + j] = r_s[j];
            if (a.xmmse) a.xmmse[f * N + j] = xh_s[j];
            if (a.var) a.var[f * N + j] = var_s[j];
        }
        if (a.traj) {
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
            block_loss<float2>(g, al, f, r_s, xh_s, a.io, t_done, bc, flags);      // Loss(T.r, T.xmmse) (vamp2.py:128)
        } else if (tid == 0) {
            bc->c[C_FRAMES] += 1;
            bc->c[C_ITERS] += t_done;
        }
        __syncthreads();
    }
    __syncthreads();
    if (a.io.counters) counters_flush(bc, a.io.counters);
}

}  // namespace

int launch_vamp2(const VampArgs& args, bool exp64, cudaStream_t stream) {
    int dev = 0, sms = 0, smem_max = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    VampArgs a = args;
    const Geom& g = a.g;
    if (3 * (size_t)g.N < (size_t)g.R) {
        set_error("vamp2: needs R <= 3 N");
        return AMPSM_ENOFIT;
    }
    const size_t vbytes = (size_t)g.R * g.N * 8;
    const bool tma_ok = vbytes % 16 == 0 && (reinterpret_cast<uintptr_t>(a.Vh) % 16) == 0 && ((size_t)a.Vh_stride * 8) % 16 == 0 &&
                        vbytes < (1u << 20);
    Vamp2Plan plan = vamp2_plan(g, tma_ok, exp64);
    a.stage_Vh = tma_ok && plan.total <= (size_t)smem_max;
    if (!a.stage_Vh) plan = vamp2_plan(g, false, exp64);
    if (plan.total > (size_t)smem_max) {
        set_error("vamp2 kernel: per-frame vectors need %zu B of shared memory (> %d B)", plan.total, smem_max);
        return AMPSM_ENOFIT;
    }
    const int threads = g.N >= 128 ? 256 : (g.N >= 64 ? 128 : ((long long)g.R * g.N <= 64 ? 32 : 64));
    auto kern = exp64 ? vamp2_kernel<true> : vamp2_kernel<false>;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total), "cudaFuncSetAttribute(vamp2)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, plan.total);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > a.frames) grid = a.frames;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, plan.total, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "vamp2_kernel launch");
}

}  // namespace ampsm
