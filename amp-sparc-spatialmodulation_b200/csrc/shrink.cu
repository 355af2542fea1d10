// Shrink family: element-wise denoisers of the reference's `random`-mode VAMP variant (shrink.py:58-157), complex64 /
// float32 arithmetic as the reference (its symbols are cast to complex64, shrink.py:26).  HBM-bound streaming kernels:
// 12 B in, 8-12 B out per entry; grid = a multiple of the SM count, grid-stride loops, coalesced 8-/4-byte accesses.
#include "kernels.h"

namespace ampsm {

// shrink.py:163-166 (regularize_exp): a[a >= log(finfo.max)] = log(finfo.max) - 1, evaluated in float32
__device__ __forceinline__ float regularize_exp(float a) {
    constexpr float kMax = 88.72283905206835f;
    return a >= kMax ? kMax - 1.0f : a;
}

// shrink.py:77-95 ('bayes'): Ps sum_k s_k G_k / (P0 G_0 + Ps sum_k G_k), G_s = exp(-|r - s|^2 / cov)
__global__ void __launch_bounds__(256) shrink_bayes_kernel(const __grid_constant__ ShrinkArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.elems; i += stride) {
        const float2 r = a.r[i];
        const float c = a.cov[a.cov_stride ? i : 0];
        float h = hypotf(r.x, r.y);                       // torch.abs(complex64) ** 2
        const float g0 = expf(-(h * h) / c);
        float gs = 0.f, sx = 0.f, sy = 0.f;
#pragma unroll 4
        for (int k = 0; k < a.al.K; ++k) {
            h = hypotf(r.x - a.al.ref[k], r.y - a.al.imf[k]);
            const float g = expf(-(h * h) / c);
            gs += g;
            sx = fmaf(a.al.ref[k], g, sx);
            sy = fmaf(a.al.imf[k], g, sy);
        }
        float norm = a.P0 * g0 + a.Ps * gs;
        if (norm == 0.f) norm = 1.0e-9f;                  // regularize_zero, shrink.py:159-161
        a.out_c[i] = make_float2(a.Ps * sx / norm, a.Ps * sy / norm);
    }
}

// shrink.py:139-157 ('shrinkOOK'): exp = 1 / (1 + eta + tol), der = nan_to_num(2 eta exp^2 / cov); the caller divides
// the accumulated sum of der by the number of entries (dxdr = der.mean()).
__global__ void __launch_bounds__(256) shrink_ook_kernel(const __grid_constant__ ShrinkArgs a) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    const float theta = logf(a.P0 / a.Ps);
    double acc = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.elems; i += stride) {
        const float c = a.cov[a.cov_stride ? i : 0];
        const float eta = expf(regularize_exp(theta + (1.0f - 2.0f * a.r[i].x) / c));
        const float e = 1.0f / (1.0f + eta + 1.0e-9f);
        float der = 2.0f * eta * e * e / c;
        if (der != der) der = 0.f;                        // torch.nan_to_num(der, nan=0.0) also maps +-inf to +-FLT_MAX
        der = fminf(fmaxf(der, -3.402823466e+38f), 3.402823466e+38f);
        a.out_f[i] = e;
        acc += (double)der;
    }
    acc = warp_sum(acc);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0 && a.sum) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += part[w];
        atomicAdd(a.sum, t);
    }
}

// shrink.py:58-75 (sw_shrinkOOK): one warp per section of M entries.
__global__ void __launch_bounds__(256) shrink_sw_ook_kernel(const __grid_constant__ ShrinkArgs a) {
    const int lane = threadIdx.x & 31;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long sections = a.elems / a.M;
    for (long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; s < sections; s += warps) {
        const long long base = s * a.M;
        float tot = 0.f;
        for (int m = lane; m < a.M; m += 32) {
            const float c = a.cov[a.cov_stride ? base + m : 0];
            tot += expf(regularize_exp((2.0f * a.r[base + m].x - 1.0f) / c));
        }
        tot = warp_sum(tot);
        for (int m = lane; m < a.M; m += 32) {
            const float c = a.cov[a.cov_stride ? base + m : 0];
            const float lr = regularize_exp((2.0f * a.r[base + m].x - 1.0f) / c);
            const float le = -logf(tot - expf(lr));
            const float eta = expf(regularize_exp(lr + le));
            const float e = eta / (1.0f + eta);
            a.out_c[base + m] = make_float2(e, 0.f);
            a.out_f[base + m] = e * (1.0f - e);
        }
    }
}

int launch_shrink(const ShrinkArgs& a, int kind, cudaStream_t stream) {
    if (a.elems <= 0) return 0;
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long threads = kind == 2 ? (a.elems / a.M) * 32 : a.elems;
    long long grid = (threads + 255) / 256;
    if (grid > (long long)sms * 8) grid = (long long)sms * 8;
    if (kind == 0) shrink_bayes_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
    else if (kind == 1) shrink_ook_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
    else shrink_sw_ook_kernel<<<(unsigned)grid, 256, 0, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "shrink kernel launch");
}

}  // namespace ampsm
