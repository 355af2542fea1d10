// Stand-alone Loss kernel: hard decision + error counters for estimates already in device memory
// (loss.py:67-179).  Persistent CTAs stride over frames; counters are flushed once per CTA.
#include "blockops.cuh"
#include "kernels.h"

namespace ampsm {

__global__ void __launch_bounds__(128) loss_kernel(const __grid_constant__ LossArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    BlockCounters* bc = reinterpret_cast<BlockCounters*>(smem);
    int* flags = reinterpret_cast<int*>(smem + sizeof(BlockCounters));
    counters_reset(bc);
    __syncthreads();
    for (long long f = blockIdx.x; f < a.frames; f += gridDim.x) {
        const int it = a.iters ? a.iters[f] : 0;
        if (a.io.x_true) {
            block_loss<float2>(a.g, a.al, f, a.xmap + f * a.g.N, a.xmmse + f * a.g.N, a.io, it, bc, flags);
        } else if (threadIdx.x == 0) {
            bc->c[C_FRAMES] += 1;
            bc->c[C_ITERS] += it;
        }
    }
    __syncthreads();
    if (a.io.counters) counters_flush(bc, a.io.counters);
}

int launch_loss(const LossArgs& a, cudaStream_t stream) {
    if (a.frames <= 0) return 0;
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t smem = sizeof(BlockCounters) + (size_t)(1 + a.g.Lin) * sizeof(int) + 16;
    long long grid = (long long)sms * 8;
    if (grid > a.frames) grid = a.frames;
    loss_kernel<<<(unsigned)grid, 128, smem, stream>>>(a);
    count_launch();
    return check_cuda(cudaGetLastError(), "loss_kernel launch");
}

}  // namespace ampsm
