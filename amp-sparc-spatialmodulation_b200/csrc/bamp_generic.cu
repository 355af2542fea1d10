// Generic BAMP kernel: one CTA per frame (persistent over frames), any n x N whose vectors fit shared memory.
// The frame's channel matrix is staged into shared memory once with a 1-D bulk TMA copy (cp.async.bulk +
// mbarrier) and all iterations run in-kernel; matrices that do not fit are read through L2 instead.
// Follows bamp.py:12-25 (state), 59-64 (iteration), 66-77 (denoiser), 116-143 (loop, exit, Loss).
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

struct BampPlan {
    size_t H, y, z, g, u, w, xh, xh_new, xmap, var, var_new, cov, scr, red, flags, bc, mbar, total;
};

__host__ __device__ inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

__host__ __device__ inline BampPlan bamp_plan(const Geom& g, bool stage, bool exp64) {
    BampPlan p;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = align16(o + bytes);
        return at;
    };
    p.H = take(stage ? (size_t)g.n * g.N * 8 : 0);
    p.y = take((size_t)g.n * 8);
    p.z = take((size_t)g.n * 8);
    p.g = take((size_t)g.n * 8);
    p.u = take((size_t)g.n * 4);
    p.w = take((size_t)g.n * 4);
    p.xh = take((size_t)g.N * 8);
    p.xh_new = take((size_t)g.N * 8);
    p.xmap = take((size_t)g.N * 8);
    p.var = take((size_t)g.N * 4);
    p.var_new = take((size_t)g.N * 4);
    p.cov = take((size_t)g.N * 4);
    p.scr = take((size_t)g.N * 3 * (exp64 ? 8 : 4));
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

template <bool EXP64>
__global__ void __launch_bounds__(256) bamp_generic_kernel(const __grid_constant__ BampArgs a) {
    using E = typename ExpT<EXP64>::type;
    extern __shared__ __align__(16) unsigned char smem[];
    const Geom& g = a.g;
    const DevAlphabet& al = a.al;
    const bool stage = a.stage_H != 0;
    const BampPlan P = bamp_plan(g, stage, EXP64);
    float2* Hs = reinterpret_cast<float2*>(smem + P.H);
    float2* y_s = reinterpret_cast<float2*>(smem + P.y);
    float2* z_s = reinterpret_cast<float2*>(smem + P.z);
    float2* g_s = reinterpret_cast<float2*>(smem + P.g);
    float* u_s = reinterpret_cast<float*>(smem + P.u);
    float* w_s = reinterpret_cast<float*>(smem + P.w);
    float2* xh_s = reinterpret_cast<float2*>(smem + P.xh);
    float2* xhn_s = reinterpret_cast<float2*>(smem + P.xh_new);
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

    counters_reset(bc);
    if (stage && tid == 0) {
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
            if (tid == 0) {
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
        if (stage && !(shared_H && H_loaded)) {
            mbar_wait(mbar, phase);
            phase ^= 1u;
            H_loaded = true;
        }
        __syncthreads();

        int t_done = 0;
        for (int t = 0; t < g.max_iters; ++t) {
            // ---- row pass: v = |H|^2 var, Hx = H xhat; then z, u and the column pass operands (bamp.py:59-61)
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
                if (lane == 0) {
                    const float2 yv = y_s[i], zo = z_s[i];
                    const float2 resid = make_float2(yv.x - zo.x, yv.y - zo.y);
                    const float2 corr = cdiv_real(make_float2(av * resid.x, av * resid.y), u_s[i]);   // old u
                    const float2 zn = make_float2(ar - corr.x, ai - corr.y);
                    const float un = av + sigma2;
                    z_s[i] = zn;
                    u_s[i] = un;
                    g_s[i] = cdiv_real(make_float2(yv.x - zn.x, yv.y - zn.y), un);
                    w_s[i] = __frcp_rn(un);
                }
            }
            __syncthreads();
            // ---- column pass: cov = 1/(|H|^2^T (1/u)), xmap = xhat + cov * H^H((y-z)/u) (bamp.py:62-63)
            for (int j = tid; j < N; j += blockDim.x) {
                float ac = 0.f, ar = 0.f, ai = 0.f;
                for (int i = 0; i < n; ++i) {
                    const float2 h = Hm[(size_t)i * N + j];
                    const float2 gv = g_s[i];
                    ac = fmaf(fmaf(h.x, h.x, h.y * h.y), w_s[i], ac);
                    ar = fmaf(h.x, gv.x, fmaf(h.y, gv.y, ar));
                    ai = fmaf(h.x, gv.y, fmaf(-h.y, gv.x, ai));
                }
                const float cov = __frcp_rn(ac);
                const float2 x = xh_s[j];
                xmap_s[j] = make_float2(fmaf(cov, ar, x.x), fmaf(cov, ai, x.y));
                cov_s[j] = cov;
            }
            __syncthreads();
            // ---- denoiser (bamp.py:66-77): tau = cov/2
            if (g.decision == 2) {                                   // generator_mode 'random' (bamp.py:46): i.i.d. prior
                block_denoise_iid(g, al, xmap_s, cov_s, xhn_s, varn_s);
            } else {
                double gshift = 0.0;
                if (EXP64 && g.shift_mode == 1) gshift = block_absmax_exponent(g, al, xmap_s, cov_s, 0.f, true, red);
                block_denoise<EXP64>(g, al, xmap_s, cov_s, 0.f, true, gshift, xhn_s, varn_s, scr);
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
                        const float2 xt = a.io.x_true[f * N + j], xe = xhn_s[j];
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
            for (int j = tid; j < N; j += blockDim.x) {
                xh_s[j] = xhn_s[j];
                var_s[j] = varn_s[j];
            }
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
    // bulk TMA needs 16-byte aligned, 16-byte multiple transfers
    const bool tma_ok = ((size_t)g.n * g.N * 8) % 16 == 0 && (reinterpret_cast<uintptr_t>(a.H) % 16) == 0 &&
                        ((size_t)a.H_stride * 8) % 16 == 0 && (size_t)g.n * g.N * 8 < (1u << 20);
    BampPlan plan = bamp_plan(g, tma_ok, exp64);
    a.stage_H = tma_ok && plan.total <= (size_t)smem_max;
    if (!a.stage_H) plan = bamp_plan(g, false, exp64);
    if (plan.total > (size_t)smem_max) {
        set_error("BAMP generic kernel: per-frame vectors need %zu B of shared memory (> %d B)", plan.total, smem_max);
        return AMPSM_ENOFIT;
    }
    const int threads = g.N >= 128 ? 256 : (g.N >= 64 ? 128 : ((long long)g.n * g.N <= 64 ? 32 : 64));
    auto kern = exp64 ? bamp_generic_kernel<true> : bamp_generic_kernel<false>;
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
