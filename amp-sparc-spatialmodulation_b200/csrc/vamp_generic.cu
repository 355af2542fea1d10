// Generic VAMP kernel: one CTA per frame (persistent over frames).  Vh is staged into shared memory once per
// frame with a 1-D bulk TMA copy when it fits; U is read once (y_tilde) straight from global memory.
// Follows vamp.py:12-28 (state), 66-94 (iteration), 96-119 (denoiser), 159-191 (loop, exit, Loss).
// CT = float2: the complex64 path; CT = double2: the reference fed with complex128 factors -- the linear stage runs
// in float64 while xmmse / var are still rounded to complex64 / float32 every iteration (vamp.py:119).
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

struct VampPlan {
    size_t Vh, yt, d, s2, rt, r, xh, var, var_new, scr, red, flags, bc, mbar, total;
};
__host__ __device__ inline size_t valign16(size_t v) { return (v + 15) & ~size_t(15); }

__host__ __device__ inline VampPlan vamp_plan(const Geom& g, bool stage, bool dbl, bool exp64) {
    VampPlan p;
    size_t o = 0;
    const size_t cs = dbl ? 16 : 8, rs = dbl ? 8 : 4;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = valign16(o + bytes);
        return at;
    };
    p.Vh = take(stage ? (size_t)g.R * g.N * cs : 0);
    p.yt = take((size_t)g.R * cs);
    p.d = take((size_t)g.R * cs);
    p.s2 = take((size_t)g.R * rs);
    p.rt = take((size_t)g.N * cs);
    p.r = take((size_t)g.N * cs);
    p.xh = take((size_t)g.N * 8);
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

template <typename RT>
struct Cplx;
template <>
struct Cplx<float> {
    using type = float2;
    static __device__ __forceinline__ float2 make(float a, float b) { return make_float2(a, b); }
};
template <>
struct Cplx<double> {
    using type = double2;
    static __device__ __forceinline__ double2 make(double a, double b) { return make_double2(a, b); }
};
__device__ __forceinline__ float rcp_rn(float x) { return __frcp_rn(x); }
__device__ __forceinline__ double rcp_rn(double x) { return __drcp_rn(x); }
__device__ __forceinline__ float clampT(float v, float lo, float hi) {
    // torch.max / torch.min propagate NaN (vamp.py:76-77)
    if (v != v) return v;
    return fminf(fmaxf(v, lo), hi);
}
__device__ __forceinline__ double clampT(double v, double lo, double hi) {
    if (v != v) return v;
    return fmin(fmax(v, lo), hi);
}

// block-wide sum of one double; valid in every thread
__device__ inline double block_sum(double a, double* red) {
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

template <typename RT, bool EXP64>
__global__ void __launch_bounds__(256) vamp_generic_kernel(const __grid_constant__ VampArgs a) {
    using CT = typename Cplx<RT>::type;
    using E = typename ExpT<EXP64>::type;
    constexpr bool DBL = sizeof(RT) == 8;
    extern __shared__ __align__(16) unsigned char smem[];
    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    const bool stage = a.stage_Vh != 0;
    const VampPlan P = vamp_plan(g, stage, DBL, EXP64);
    CT* Vs = reinterpret_cast<CT*>(smem + P.Vh);
    CT* yt_s = reinterpret_cast<CT*>(smem + P.yt);
    CT* d_s = reinterpret_cast<CT*>(smem + P.d);
    RT* s2_s = reinterpret_cast<RT*>(smem + P.s2);
    CT* rt_s = reinterpret_cast<CT*>(smem + P.rt);
    CT* r_s = reinterpret_cast<CT*>(smem + P.r);
    float2* xh_s = reinterpret_cast<float2*>(smem + P.xh);
    float* var_s = reinterpret_cast<float*>(smem + P.var);
    float* varn_s = reinterpret_cast<float*>(smem + P.var_new);
    E* scr = reinterpret_cast<E*>(smem + P.scr);
    double* red = reinterpret_cast<double*>(smem + P.red);
    int* flags = reinterpret_cast<int*>(smem + P.flags);
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem + P.bc);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + P.mbar);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const int n = g.n, N = g.N, R = g.R;
    const uint32_t Vbytes = (uint32_t)((size_t)R * N * sizeof(CT));
    const bool shared_V = a.Vh_stride == 0;
    const RT ratio_min = (RT)1.0e-5f, ratio_max = (RT)(1.0f - 1.0e-5f);   // float32 tensors (vamp.py:51-52)
    const RT var_min = (RT)1.0e-9f, var_max = (RT)1.0e5f;                 // vamp.py:53-54
    const double eta_d = (double)R / (double)N;                            // vamp.py:28
    const RT eta = (RT)eta_d, one_m_eta = (RT)(1.0 - eta_d);

    counters_reset(bc);
    if (stage && tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    bool V_loaded = false;

    for (long long f = blockIdx.x; f < a.frames; f += gridDim.x) {
        const CT* Vg = reinterpret_cast<const CT*>(a.Vh) + f * a.Vh_stride;
        const CT* Vm = stage ? Vs : Vg;
        const CT* Ug = reinterpret_cast<const CT*>(a.U) + f * a.U_stride;
        const RT* sg = reinterpret_cast<const RT*>(a.s) + f * a.s_stride;
        const CT* yg = reinterpret_cast<const CT*>(a.y) + f * n;
        if (stage && !(shared_V && V_loaded) && tid == 0) {
            mbar_expect_tx(mbar, Vbytes);
            tma_load_1d(Vs, Vg, Vbytes, mbar);
        }
        const double noise_var_d = a.sigma2_pf ? (double)a.sigma2_pf[f] : a.sigma2_d;
        const RT nv = (RT)noise_var_d;
        // y_tilde = (s * U^H) y  (vamp.py:22): thread per singular value, U read column-wise (coalesced over k)
        for (int k = tid; k < R; k += blockDim.x) {
            RT ar = 0, ai = 0;
            const RT sk = sg[k];
            for (int i = 0; i < n; ++i) {
                const CT u = Ug[(size_t)i * R + k];
                const CT yv = yg[i];
                const RT wr = sk * u.x, wi = -(sk * u.y);           // s * conj(U)
                ar += wr * yv.x - wi * yv.y;
                ai += wr * yv.y + wi * yv.x;
            }
            yt_s[k] = Cplx<RT>::make(ar, ai);
            s2_s[k] = sk * sk;                                      // vamp.py:17
        }
        const double sp = a.sparsity;
        for (int j = tid; j < N; j += blockDim.x) {
            rt_s[j] = Cplx<RT>::make((RT)sp, (RT)0);                // vamp.py:25
            r_s[j] = Cplx<RT>::make((RT)0, (RT)0);
            var_s[j] = 1.0f;
            xh_s[j] = make_float2(0.f, 0.f);
        }
        double s2t_d = sp * sp * (1.0 - sp) + (1.0 - sp) * (1.0 - sp) * sp;   // python float (vamp.py:26)
        RT s2t = (RT)s2t_d;
        if (stage && !(shared_V && V_loaded)) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            V_loaded = true;
        }
        __syncthreads();

        int t_done = 0;
        for (int t = 0; t < g.max_iters; ++t) {
            // var_ratio: python-float division on the first pass, tensor division afterwards (vamp.py:66)
            const RT ratio = (t == 0) ? (RT)(noise_var_d / s2t_d) : nv / s2t;
            // ---- q = Vh r_tilde; d = scale (y_tilde + ratio q) - q   (vamp.py:67-72)
            double scale_sum = 0.0;
            for (int k = warp; k < R; k += nwarps) {
                const CT* Vrow = Vm + (size_t)k * N;
                RT ar = 0, ai = 0;
                for (int j = lane; j < N; j += 32) {
                    const CT v = Vrow[j];
                    const CT x = rt_s[j];
                    ar += v.x * x.x - v.y * x.y;
                    ai += v.x * x.y + v.y * x.x;
                }
                ar = warp_sum(ar);
                ai = warp_sum(ai);
                if (lane == 0) {
                    const RT scale = rcp_rn(s2_s[k] + ratio);
                    const CT yt = yt_s[k];
                    d_s[k] = Cplx<RT>::make(scale * (yt.x + ratio * ar) - ar, scale * (yt.y + ratio * ai) - ai);
                    scale_sum += (double)scale;
                }
            }
            const double scale_tot = block_sum(scale_sum, red);     // also orders d_s before the column pass
            const RT var_lmmse = (RT)(scale_tot / R) * nv;          // vamp.py:71
            const RT xt_var = eta * var_lmmse + one_m_eta * s2t;    // vamp.py:73
            const RT alpha = clampT(xt_var / s2t, ratio_min, ratio_max);
            const RT inv_1ma = rcp_rn((RT)1 - alpha);
            const RT sig2 = clampT(alpha / ((RT)1 - alpha) * s2t, var_min, var_max);   // vamp.py:80-82
            // ---- x_tilde = V (.) + r_tilde ; r = (x_tilde - alpha r_tilde) / (1 - alpha)  (vamp.py:72,79)
            for (int j = tid; j < N; j += blockDim.x) {
                RT ar = 0, ai = 0;
                for (int k = 0; k < R; ++k) {
                    const CT v = Vm[(size_t)k * N + j];
                    const CT dv = d_s[k];
                    ar += v.x * dv.x + v.y * dv.y;                  // conj(v) * d
                    ai += v.x * dv.y - v.y * dv.x;
                }
                const CT rt = rt_s[j];
                const RT xr = ar + rt.x, xi = ai + rt.y;
                r_s[j] = Cplx<RT>::make((xr - alpha * rt.x) * inv_1ma, (xi - alpha * rt.y) * inv_1ma);
            }
            __syncthreads();
            // ---- denoiser with the scalar, un-halved variance (vamp.py:84, 96-119)
            double gshift = 0.0;
            if (EXP64 && g.shift_mode == 1) gshift = block_absmax_exponent<CT>(g, al, r_s, nullptr, sig2, false, red);
            block_denoise<EXP64, CT>(g, al, r_s, nullptr, sig2, false, gshift, xh_s, varn_s, scr);
            __syncthreads();
            // ---- Onsager bookkeeping (vamp.py:85-94) and the exit test on var (vamp.py:185)
            double vsum = 0.0;
            bool close = true;
            for (int j = tid; j < N; j += blockDim.x) {
                const float vn = varn_s[j], vo = var_s[j];
                vsum += vn;
                close &= fabsf(vn - vo) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, vo)));
            }
            const int all_close = __syncthreads_and(close ? 1 : 0);
            const double vtot = block_sum(vsum, red);
            const RT dxdr = clampT((RT)(float)(vtot / N) / sig2, ratio_min, ratio_max);
            const RT norm = rcp_rn((RT)1 - dxdr);
            double mse = 0.0;
            for (int j = tid; j < N; j += blockDim.x) {
                const float2 xe = xh_s[j];
                const CT rv = r_s[j];
                rt_s[j] = Cplx<RT>::make(((RT)xe.x - dxdr * rv.x) * norm, ((RT)xe.y - dxdr * rv.y) * norm);
                var_s[j] = varn_s[j];
                if (a.traj && a.io.x_true) {
                    const float2 xt = a.io.x_true[f * N + j];
                    const double dr = (double)xe.x - xt.x, di = (double)xe.y - xt.y;
                    mse += dr * dr + di * di;
                }
            }
            s2t = clampT(sig2 * dxdr * norm, var_min, var_max);
            if (a.traj) {
                mse = block_sum(mse, red);
                if (tid == 0) {
                    float* tr = a.traj + (f * g.max_iters + t) * 3;
                    tr[0] = (float)s2t;
                    tr[1] = (float)(vtot / N);
                    tr[2] = (float)(mse / N);
                }
            }
            __syncthreads();
            t_done = t + 1;
            if (g.early_exit && all_close) break;
        }
        for (int j = tid; j < N; j += blockDim.x) {
            if (a.xmap) reinterpret_cast<CT*>(a.xmap)[f * N + j] = r_s[j];
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
            block_loss<CT>(g, al, f, r_s, xh_s, a.io, t_done, bc, flags);   // Loss is fed T.r as xmap (vamp.py:187)
        } else if (tid == 0) {
            bc->c[C_FRAMES] += 1;
            bc->c[C_ITERS] += t_done;
        }
        __syncthreads();
    }
    __syncthreads();
    if (a.io.counters) counters_flush(bc, a.io.counters);
}

int launch_vamp_generic(const VampArgs& args, bool is_double, bool exp64, cudaStream_t stream) {
    int dev = 0, sms = 0, smem_max = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    VampArgs a = args;
    const Geom& g = a.g;
    if (is_double) exp64 = true;
    const size_t cs = is_double ? 16 : 8;
    const size_t vbytes = (size_t)g.R * g.N * cs;
    const bool tma_ok = vbytes % 16 == 0 && (reinterpret_cast<uintptr_t>(a.Vh) % 16) == 0 &&
                        ((size_t)a.Vh_stride * cs) % 16 == 0 && vbytes < (1u << 20);
    VampPlan plan = vamp_plan(g, tma_ok, is_double, exp64);
    a.stage_Vh = tma_ok && plan.total <= (size_t)smem_max;
    if (!a.stage_Vh) plan = vamp_plan(g, false, is_double, exp64);
    if (plan.total > (size_t)smem_max) {
        set_error("VAMP generic kernel: per-frame vectors need %zu B of shared memory (> %d B)", plan.total, smem_max);
        return AMPSM_ENOFIT;
    }
    const int threads = g.N >= 128 ? 256 : (g.N >= 64 ? 128 : ((long long)g.R * g.N <= 64 ? 32 : 64));
    void (*kern)(const VampArgs);
    if (is_double)
        kern = vamp_generic_kernel<double, true>;
    else
        kern = exp64 ? vamp_generic_kernel<float, true> : vamp_generic_kernel<float, false>;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total),
                           "cudaFuncSetAttribute(vamp_generic)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, plan.total);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > a.frames) grid = a.frames;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, threads, plan.total, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "vamp_generic_kernel launch");
}

}  // namespace ampsm
