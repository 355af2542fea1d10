// Micro-benchmark: how fast can ONE warp (or two) per SM sub-partition issue FFMA2 / FFMA streams shaped like the BAMP
// mat-vec passes (a register-resident tile times broadcast operands into a few accumulator chains)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_issue ffma2_issue.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long pair_t;
__device__ __forceinline__ pair_t ffma2(pair_t a, pair_t b, pair_t c) {
    pair_t d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// MODE 0: operand x changes every instruction (no operand reuse)          192 FFMA2, 12 chains
// MODE 1: x constant over the whole stream                                192 FFMA2, 12 chains
// MODE 2: x reused by 4 consecutive instructions (row pass shape)         192 FFMA2, 12 chains
// MODE 3: tile operand reused by 2 consecutive instructions (A and B of the row pass: same h, x0.x / x0.y)
// MODE 4: scalar FFMA, x reused by 8 consecutive                          384 FFMA, 24 chains
// MODE 5: scalar FFMA, no reuse
template <int MODE>
__global__ void k(const pair_t* in, pair_t* out, long long* cyc, int reps) {
    pair_t H[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) H[i] = in[threadIdx.x + 32 * i];
    pair_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = in[threadIdx.x + 7 + i];
    pair_t acc[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) acc[i] = 0ull;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll
            for (int t = 0; t < 16; ++t)
#pragma unroll
                for (int c = 0; c < 12; ++c) acc[c] = ffma2(H[(t * 12 + c) & 63], x[(t + c) & 7], acc[c]);
        } else if (MODE == 1) {
#pragma unroll
            for (int t = 0; t < 16; ++t)
#pragma unroll
                for (int c = 0; c < 12; ++c) acc[c] = ffma2(H[(t * 12 + c) & 63], x[0], acc[c]);
        } else if (MODE == 2) {
#pragma unroll
            for (int t = 0; t < 48; ++t)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[(t % 3) * 4 + c] = ffma2(H[(t * 4 + c) & 63], x[t & 7], acc[(t % 3) * 4 + c]);
        } else if (MODE == 3) {
#pragma unroll
            for (int t = 0; t < 16; ++t)
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    acc[2 * c] = ffma2(H[(t * 6 + c) & 63], x[(2 * t) & 7], acc[2 * c]);
                    acc[2 * c + 1] = ffma2(H[(t * 6 + c) & 63], x[(2 * t + 1) & 7], acc[2 * c + 1]);
                }
        } else {
            float* hf = reinterpret_cast<float*>(H);
            float* xf = reinterpret_cast<float*>(x);
            float* af = reinterpret_cast<float*>(acc);
#pragma unroll
            for (int t = 0; t < 16; ++t)
#pragma unroll
                for (int c = 0; c < 24; ++c)
                    af[c] = fmaf(hf[(t * 24 + c) & 127], MODE == 4 ? xf[(3 * t + c / 8) & 15] : xf[(t + c) & 15], af[c]);
        }
    }
    long long t1 = clock64();
    pair_t s = 0;
#pragma unroll
    for (int i = 0; i < 12; ++i) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps_per_sm, int per_rep) {
    pair_t *in, *out;
    long long* cyc;
    cudaMalloc(&in, 1 << 20);
    cudaMemset(in, 0x3c, 1 << 20);
    cudaMalloc(&out, 1 << 24);
    cudaMalloc(&cyc, 8);
    const int reps = 2000;
    for (int it = 0; it < 2; ++it) k<MODE><<<148, 32 * warps_per_sm>>>(in, out, cyc, reps);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s warps/SM=%2d  cycles per rep = %8.1f  (%.2f cycles/instr/warp)  err=%s\n", name, warps_per_sm,
           (double)h / reps, (double)h / reps / per_rep, cudaGetErrorString(cudaGetLastError()));
    cudaFree(in); cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8}) run<0>("FFMA2 x192 no reuse", w, 192);
    for (int w : {4, 8}) run<1>("FFMA2 x192 x constant", w, 192);
    for (int w : {4, 8}) run<2>("FFMA2 x192 x reused by 4", w, 192);
    for (int w : {4, 8}) run<3>("FFMA2 x192 h reused by 2", w, 192);
    for (int w : {4, 8}) run<4>("FFMA x384 x reused by 8", w, 384);
    for (int w : {4, 8}) run<5>("FFMA x384 no reuse", w, 384);
    return 0;
}
