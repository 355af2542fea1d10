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
    unsigned char* bplanes0;   // structured path (scamp_st.cu): pre-split taps of the residual GEMM
    unsigned char* bplanes1;   // ... and of the estimate GEMM
};

// structured (taps) tensor-core path, scamp_st.cu
struct ScampStPlan {
    bool ok;
    int brows0, nks0;            // residual: MMA N (Lh Nr padded to 32), K stages
    int brows1, chunks1, nks1;   // estimate: outputs per chunk, chunks per column block, K stages
    int kpb;                     // estimate: K stages per tap
    int FR;                      // frames per tile
    int zs_stride;               // complex elements per frame of the padded Zs
    size_t bplane_bytes0, bplane_bytes1;
};
ScampStPlan scamp_st_plan(const Geom& g, int Lh);
int scamp_st_prepare(const Geom& g, const ScampStPlan& p, int Lh, const float2* taps, unsigned char* bplanes0, unsigned char* bplanes1,
                     cudaStream_t stream);
bool scamp_st_can_fuse(const Geom& g, const DevAlphabet& al);   // section denoiser + psi + exit test in the estimate GEMM's epilogue
int scamp_st_gemm(int mode, const ScampWs& w, const Geom& g, const ScampStPlan& p, int Lh, const unsigned char* bplanes, const float2* y,
                  long long F, const DevAlphabet& al, bool fused, const float* W, float sigma2, const float* sigma2_pf, int t, cudaStream_t stream);
// the residual mode computes the block scalars (gamma, b, phi, tau) of its frames itself; the fused estimate mode retires its frames

// tensor-core GEMMs (scamp_tc.cu): mode 0 = residual with Bm = A, mode 1 = estimate with Bm = A^T
bool scamp_tc_fits(int n, int N, long long F);
int scamp_tc_prepare(const float2* A, float2* At, int n, int N, cudaStream_t stream);
int scamp_tc_gemm(int mode, const ScampWs& w, const Geom& g, const float2* Bm, const float2* y, long long F, cudaStream_t stream);

}  // namespace ampsm
