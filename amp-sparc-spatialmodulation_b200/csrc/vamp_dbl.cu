// Register-resident VAMP kernel for complex128 factors at BASELINE config 3 (Vh 64 x 128 complex128 = 128 KiB): the
// reference fed with upcast inputs (vamp.py:12-28; linear stage in float64, denoiser outputs rounded to complex64 / float32
// every iteration, vamp.py:119) on the FP64 pipe.
//
// One CTA of 256 threads per frame (persistent over frames), Vh in REGISTERS for all iterations: warp w keeps rows
// 8 w .. 8 w + 7, its lanes form a 2 x 16 grid, lane (a, b) holding the 4 x 8 tile of rows 8 w + 4 a + i, columns b + 16 t
// (32 complex128 = 128 registers).  Both mat-vecs of an iteration run from these registers with DFMA:
//   row pass    q = Vh r~      : 4 x 8 complex MACs per lane, the 16 column groups of a row reduced by 4 shuffle rounds;
//   column pass x~ = V d + r~  : 8 x 4 conjugate MACs per lane, the 2 row groups by one shuffle round, the 8 warps through
//                                a 16 KiB shared-memory plane.
// The generic kernel (csrc/vamp_generic.cu) reads the 128 KiB matrix from shared memory twice per iteration with 8 warps
// per SM: 1.8 TFLOP/s, 5 % of the DFMA peak (round 2 bench).  Everything outside the two mat-vecs -- the LMMSE scalars with
// the reference's clips, the float64 denoiser (block_denoise), the exit test, the fused Loss -- is the generic kernel's code,
// so the results are the generic kernel's to summation order.
// launch_vamp_dbl() returns AMPSM_ENOFIT for other shapes.
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

namespace {

constexpr int kR = 64, kN = 128, kThreads = 256, kWarps = 8;
constexpr int RT_ = 4, CT_ = 8;       // tile rows / columns per lane

struct DblPlan {
    size_t yt, d, s2, rt, r, xh, var, var_new, scr, red, flags, bc, colp, total;
};
__host__ __device__ inline size_t dalign16(size_t v) { return (v + 15) & ~size_t(15); }
__host__ __device__ inline DblPlan dbl_plan(const Geom& g) {
    DblPlan p;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = dalign16(o + bytes);
        return at;
    };
    p.yt = take((size_t)kR * 16);
    p.d = take((size_t)kR * 16);
    p.s2 = take((size_t)kR * 8);
    p.rt = take((size_t)kN * 16);
    p.r = take((size_t)kN * 16);
    p.xh = take((size_t)kN * 8);
    p.var = take((size_t)kN * 4);
    p.var_new = take((size_t)kN * 4);
    p.scr = take((size_t)kN * 3 * 8);
    p.red = take(32 * 3 * 8);
    p.flags = take((size_t)(1 + g.Lin) * 4);
    p.bc = take(sizeof(BlockCounters));
    p.colp = take((size_t)kWarps * kN * 16);          // column partials of the 8 warps
    p.total = o;
    return p;
}

__device__ __forceinline__ double clampD(double v, double lo, double hi) {
    if (v != v) return v;                              // torch.max / torch.min propagate NaN (vamp.py:76-77)
    return fmin(fmax(v, lo), hi);
}
__device__ inline double block_sum_d(double a, double* red) {
    a = warp_sum(a);
    const int warp = threadIdx.x >> 5;
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[warp] = a;
    __syncthreads();
    double r = 0.0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) r += red[w];
    __syncthreads();
    return r;
}
__device__ __forceinline__ double shfl_xor_d(double v, int o) { return __shfl_xor_sync(0xffffffffu, v, o); }

__global__ void __launch_bounds__(kThreads, 1) vamp_dbl_kernel(const __grid_constant__ VampArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    const DblPlan P = dbl_plan(g);
    double2* yt_s = reinterpret_cast<double2*>(smem + P.yt);
    double2* d_s = reinterpret_cast<double2*>(smem + P.d);
    double* s2_s = reinterpret_cast<double*>(smem + P.s2);
    double2* rt_s = reinterpret_cast<double2*>(smem + P.rt);
    double2* r_s = reinterpret_cast<double2*>(smem + P.r);
    float2* xh_s = reinterpret_cast<float2*>(smem + P.xh);
    float* var_s = reinterpret_cast<float*>(smem + P.var);
    float* varn_s = reinterpret_cast<float*>(smem + P.var_new);
    double* scr = reinterpret_cast<double*>(smem + P.scr);
    double* red = reinterpret_cast<double*>(smem + P.red);
    int* flags = reinterpret_cast<int*>(smem + P.flags);
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem + P.bc);
    double2* colp = reinterpret_cast<double2*>(smem + P.colp);

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int la = lane >> 4, lb = lane & 15;
    const int row0 = 8 * w + RT_ * la;                 // first row of the lane's tile; its columns are lb + 16 t
    const int n = g.n;
    constexpr int N = kN, R = kR;
    const double ratio_min = (double)1.0e-5f, ratio_max = (double)(1.0f - 1.0e-5f);   // float32 tensors (vamp.py:51-52)
    const double var_min = (double)1.0e-9f, var_max = (double)1.0e5f;                 // vamp.py:53-54
    const double eta = (double)R / (double)N, one_m_eta = 1.0 - eta;                  // vamp.py:28

    counters_reset(bc);
    __syncthreads();
    const bool shared_V = a.Vh_stride == 0;
    double2 V[RT_][CT_];
    bool V_loaded = false;

    for (long long f = blockIdx.x; f < a.frames; f += gridDim.x) {
        const double2* Vg = reinterpret_cast<const double2*>(a.Vh) + f * a.Vh_stride;
        const double2* Ug = reinterpret_cast<const double2*>(a.U) + f * a.U_stride;
        const double* sg = reinterpret_cast<const double*>(a.s) + f * a.s_stride;
        const double2* yg = reinterpret_cast<const double2*>(a.y) + f * n;
        if (!(shared_V && V_loaded)) {                 // the tile: per (i, t) a half-warp reads 16 consecutive complex128 (256 bytes)
#pragma unroll
            for (int i = 0; i < RT_; ++i)
#pragma unroll
                for (int t = 0; t < CT_; ++t) V[i][t] = __ldg(Vg + (size_t)(row0 + i) * N + lb + 16 * t);
            V_loaded = true;
        }
        const double noise_var_d = a.sigma2_pf ? (double)a.sigma2_pf[f] : a.sigma2_d;
        const double nv = noise_var_d;
        // y_tilde = (s * U^H) y  (vamp.py:22): four threads per singular value (rows of U interleaved), U read column-wise
        {
            const int k = tid >> 2, part = tid & 3;
            double ar = 0, ai = 0;
            const double sk = sg[k];
            for (int i = part; i < n; i += 4) {
                const double2 u = __ldg(Ug + (size_t)i * R + k);
                const double2 yv = yg[i];
                const double wr = sk * u.x, wi = -(sk * u.y);           // s * conj(U)
                ar += wr * yv.x - wi * yv.y;
                ai += wr * yv.y + wi * yv.x;
            }
            ar += shfl_xor_d(ar, 1);
            ai += shfl_xor_d(ai, 1);
            ar += shfl_xor_d(ar, 2);
            ai += shfl_xor_d(ai, 2);
            if (part == 0) {
                yt_s[k] = make_double2(ar, ai);
                s2_s[k] = sk * sk;                                      // vamp.py:17
            }
        }
        const double sp = a.sparsity;
        if (tid < N) {
            rt_s[tid] = make_double2(sp, 0.0);                          // vamp.py:25
            r_s[tid] = make_double2(0.0, 0.0);
            var_s[tid] = 1.0f;
            xh_s[tid] = make_float2(0.f, 0.f);
        }
        const double s2t_d = sp * sp * (1.0 - sp) + (1.0 - sp) * (1.0 - sp) * sp;   // python float (vamp.py:26)
        double s2t = s2t_d;
        __syncthreads();

        int t_done = 0;
        for (int t = 0; t < g.max_iters; ++t) {
            // var_ratio: python-float division on the first pass, tensor division afterwards (vamp.py:66)
            const double ratio = (t == 0) ? (noise_var_d / s2t_d) : nv / s2t;
            // ================= row pass: q = Vh r~ (vamp.py:67), then d = scale (y~ + ratio q) - q (vamp.py:68-72) =================
            double scale_sum = 0.0;
            {
                double qr[RT_], qi[RT_];
#pragma unroll
                for (int i = 0; i < RT_; ++i) qr[i] = qi[i] = 0.0;
#pragma unroll
                for (int tt = 0; tt < CT_; ++tt) {
                    const double2 x = rt_s[lb + 16 * tt];
#pragma unroll
                    for (int i = 0; i < RT_; ++i) {
                        const double2 v = V[i][tt];
                        qr[i] = fma(v.x, x.x, fma(-v.y, x.y, qr[i]));
                        qi[i] = fma(v.x, x.y, fma(v.y, x.x, qi[i]));
                    }
                }
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) {
#pragma unroll
                    for (int i = 0; i < RT_; ++i) {
                        qr[i] += shfl_xor_d(qr[i], o);
                        qi[i] += shfl_xor_d(qi[i], o);
                    }
                }
                if (lb < RT_) {                        // lanes b = 0..3 of each half finish one row each
                    const int i = lb, row = row0 + i;
                    const double ar = i == 0 ? qr[0] : i == 1 ? qr[1] : i == 2 ? qr[2] : qr[3];
                    const double ai = i == 0 ? qi[0] : i == 1 ? qi[1] : i == 2 ? qi[2] : qi[3];
                    const double scale = __drcp_rn(s2_s[row] + ratio);
                    const double2 yt = yt_s[row];
                    d_s[row] = make_double2(scale * (yt.x + ratio * ar) - ar, scale * (yt.y + ratio * ai) - ai);
                    scale_sum = scale;
                }
            }
            const double scale_tot = block_sum_d(scale_sum, red);       // also orders d_s before the column pass
            const double var_lmmse = (scale_tot / R) * nv;              // vamp.py:71
            const double xt_var = eta * var_lmmse + one_m_eta * s2t;    // vamp.py:73
            const double alpha = clampD(xt_var / s2t, ratio_min, ratio_max);
            const double inv_1ma = __drcp_rn(1.0 - alpha);
            const double sig2 = clampD(alpha / (1.0 - alpha) * s2t, var_min, var_max);   // vamp.py:80-82
            // ================= column pass: V d (vamp.py:72), partial over the lane's 4 rows =================
            {
                double cr[CT_], ci[CT_];
#pragma unroll
                for (int tt = 0; tt < CT_; ++tt) cr[tt] = ci[tt] = 0.0;
#pragma unroll
                for (int i = 0; i < RT_; ++i) {
                    const double2 dv = d_s[row0 + i];
#pragma unroll
                    for (int tt = 0; tt < CT_; ++tt) {
                        const double2 v = V[i][tt];
                        cr[tt] = fma(v.x, dv.x, fma(v.y, dv.y, cr[tt]));      // conj(v) * d
                        ci[tt] = fma(v.x, dv.y, fma(-v.y, dv.x, ci[tt]));
                    }
                }
#pragma unroll
                for (int tt = 0; tt < CT_; ++tt) {
                    cr[tt] += shfl_xor_d(cr[tt], 16);
                    ci[tt] += shfl_xor_d(ci[tt], 16);
                }
                if (la == 0) {
#pragma unroll
                    for (int tt = 0; tt < CT_; ++tt) colp[w * N + lb + 16 * tt] = make_double2(cr[tt], ci[tt]);
                }
            }
            __syncthreads();
            // x_tilde = V (.) + r_tilde ; r = (x_tilde - alpha r_tilde) / (1 - alpha)  (vamp.py:72,79)
            if (tid < N) {
                double ar = 0, ai = 0;
#pragma unroll
                for (int ww = 0; ww < kWarps; ++ww) {
                    const double2 p = colp[ww * N + tid];
                    ar += p.x;
                    ai += p.y;
                }
                const double2 rt = rt_s[tid];
                const double xr = ar + rt.x, xi = ai + rt.y;
                r_s[tid] = make_double2((xr - alpha * rt.x) * inv_1ma, (xi - alpha * rt.y) * inv_1ma);
            }
            __syncthreads();
            // ---- denoiser with the scalar, un-halved variance (vamp.py:84, 96-119): float64 exponents
            double gshift = 0.0;
            if (g.shift_mode == 1) gshift = block_absmax_exponent<double2>(g, al, r_s, nullptr, sig2, false, red);
            block_denoise<true, double2>(g, al, r_s, nullptr, sig2, false, gshift, xh_s, varn_s, scr);
            __syncthreads();
            // ---- Onsager bookkeeping (vamp.py:85-94) and the exit test on var (vamp.py:185)
            double vsum = 0.0;
            bool close = true;
            if (tid < N) {
                const float vn = varn_s[tid], vo = var_s[tid];
                vsum = vn;
                close = fabsf(vn - vo) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, vo)));
            }
            const int all_close = __syncthreads_and(close ? 1 : 0);
            const double vtot = block_sum_d(vsum, red);
            const double dxdr = clampD((double)(float)(vtot / N) / sig2, ratio_min, ratio_max);
            const double norm = __drcp_rn(1.0 - dxdr);
            double mse = 0.0;
            if (tid < N) {
                const float2 xe = xh_s[tid];
                const double2 rv = r_s[tid];
                rt_s[tid] = make_double2(((double)xe.x - dxdr * rv.x) * norm, ((double)xe.y - dxdr * rv.y) * norm);
                var_s[tid] = varn_s[tid];
                if (a.traj && a.io.x_true) {
                    const float2 xt = a.io.x_true[f * N + tid];
                    const double dr = (double)xe.x - xt.x, di = (double)xe.y - xt.y;
                    mse = dr * dr + di * di;
                }
            }
            s2t = clampD(sig2 * dxdr * norm, var_min, var_max);
            if (a.traj) {
                mse = block_sum_d(mse, red);
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
        if (tid < N) {
            if (a.xmap) reinterpret_cast<double2*>(a.xmap)[f * N + tid] = r_s[tid];
            if (a.xmmse) a.xmmse[f * N + tid] = xh_s[tid];
            if (a.var) a.var[f * N + tid] = var_s[tid];
        }
        if (a.traj) {
            for (int t = t_done + tid; t < g.max_iters; t += kThreads) {
                const float* last = a.traj + (f * g.max_iters + t_done - 1) * 3;
                float* tr = a.traj + (f * g.max_iters + t) * 3;
                tr[0] = last[0];
                tr[1] = last[1];
                tr[2] = last[2];
            }
        }
        if (tid == 0 && a.iters) a.iters[f] = t_done;
        if (a.io.x_true) {
            block_loss<double2>(g, al, f, r_s, xh_s, a.io, t_done, bc, flags);   // Loss is fed T.r as xmap (vamp.py:187)
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

int launch_vamp_dbl(const VampArgs& args, cudaStream_t stream) {
    const Geom& g = args.g;
    if (g.R != kR || g.N != kN || g.n < 1 || g.max_iters < 1 || getenv("AMPSM_VAMP_DBL_GENERIC")) return AMPSM_ENOFIT;
    if ((reinterpret_cast<uintptr_t>(args.Vh) % 16) || (reinterpret_cast<uintptr_t>(args.U) % 16) || (reinterpret_cast<uintptr_t>(args.y) % 16))
        return AMPSM_ENOFIT;
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const DblPlan plan = dbl_plan(g);
    if (int e = check_cuda(cudaFuncSetAttribute(vamp_dbl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.total),
                           "cudaFuncSetAttribute(vamp_dbl)"))
        return e;
    long long grid = sms;
    if (grid > args.frames) grid = args.frames;
    if (grid < 1) grid = 1;
    vamp_dbl_kernel<<<(unsigned)grid, kThreads, plan.total, stream>>>(args);
    count_launch();
    return check_cuda(cudaGetLastError(), "vamp_dbl_kernel launch");
}

}  // namespace ampsm
