// SCAMP workspace shared by the SIMT path (scamp.cu) and the tensor-core GEMMs (scamp_tc.cu).
#pragma once
#include "common.cuh"

namespace ampsm {

constexpr int TILE = 32; // granularity of the zero-tile map of A

struct ScampWs {
    float2* Xh;      // [F][N]
    float2* Z;       // [F][n]
    float2* Zs;      // [F][n]   Z / phi
    float2* Xmap;    // [F][N]
    float* psi;      // [F][Lc]
    float* phi;      // [F][Lr]
    float* tau;      // [F][Lc]
    float* b;        // [F][Lr]
    int* active;     // [F]
    int* iters;      // [F]
    void* scr;       // [F][3N] exponent-typed scratch of the denoiser
    unsigned char* nz;  // [ceil(n/TILE)][ceil(N/TILE)]
    int nzc;         // columns of nz
    float2* At;      // [N][n] transpose of A (tensor-core path only)
    int* notclose;   // [F] set by the fast denoiser when a column block of the frame failed the exit test
};

// tensor-core GEMMs (scamp_tc.cu): mode 0 = residual with Bm = A, mode 1 = estimate with Bm = A^T
bool scamp_tc_fits(int n, int N, long long F);
int scamp_tc_prepare(const float2* A, float2* At, int n, int N, cudaStream_t stream);
int scamp_tc_gemm(int mode, const ScampWs& w, const Geom& g, const float2* Bm, const float2* y, long long F, cudaStream_t stream);

}  // namespace ampsm
