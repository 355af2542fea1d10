// FP32 FFMA throughput probe: the roofline denominator for the register / shared-memory resident iterations
// (MEASURED_PEAKS.json holds only the HBM copy bandwidth and the bf16 GEMM rate).
#include "kernels.h"

namespace ampsm {

__global__ void __launch_bounds__(256) ffma_probe_kernel(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

// same arithmetic issued as packed fp32x2 FMAs (FFMA2, sm_100): half the instructions for the same flops
__global__ void __launch_bounds__(256) ffma2_probe_kernel(float* out, int iters, float a, float b) {
    float2 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = make_float2(threadIdx.x * 1e-3f + u, threadIdx.x * 2e-3f + u);
    const float2 aa = make_float2(a, a), bb = make_float2(b, b);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
#pragma unroll
            for (int v = 0; v < 8; ++v) x[v] = __ffma2_rn(x[v], aa, bb);
        }
    }
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) s += x[u].x + x[u].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 DFMA throughput: the roofline denominator of the complex128 VAMP kernels
__global__ void __launch_bounds__(256) dfma_probe_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1., x2 = x0 + 2., x3 = x0 + 3., x4 = x0 + 4., x5 = x0 + 5., x6 = x0 + 6., x7 = x0 + 7.;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int probe_fp64(int device, double* tflops) {
    if (!tflops) return AMPSM_EINVAL;
    if (int e = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 1024;
    double* out = nullptr;
    if (int e = check_cuda(cudaMalloc(&out, (size_t)blocks * threads * 8), "cudaMalloc(probe)")) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dfma_probe_kernel<<<blocks, threads>>>(out, iters, 0.999, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return check_cuda(cudaGetLastError(), "dfma probe");
}

int probe_fp32x2(int device, double* tflops) {
    if (!tflops) return AMPSM_EINVAL;
    if (int e = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    if (int e = check_cuda(cudaMalloc(&out, (size_t)blocks * threads * 4), "cudaMalloc(probe)")) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        ffma2_probe_kernel<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 2 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return check_cuda(cudaGetLastError(), "ffma2 probe");
}

int probe_fp32(int device, double* tflops) {
    if (!tflops) return AMPSM_EINVAL;
    if (int e = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return e;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    float* out = nullptr;
    if (int e = check_cuda(cudaMalloc(&out, (size_t)blocks * threads * 4), "cudaMalloc(probe)")) return e;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        ffma_probe_kernel<<<blocks, threads>>>(out, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return check_cuda(cudaGetLastError(), "ffma probe");
}

}  // namespace ampsm
