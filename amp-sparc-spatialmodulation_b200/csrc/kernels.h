// Kernel argument blocks and launchers (internal; the public surface is include/ampsm_b200.h).
#pragma once
#include "common.cuh"

namespace ampsm {

constexpr int kGridMax = 4;
constexpr int kGridCorrMax = 2;
// Product-grid view of an alphabet (see bamp_fast.cu): sorted real / imaginary levels, in the log2 domain for the
// exponents (x log2 e) and plain for the moments, plus the grid points whose multiplicity differs from one.
struct DevGrid {
    int ok, nr, ni, ncorr;
    double lr2[kGridMax], li2[kGridMax];
    float lr2f[kGridMax], li2f[kGridMax], lrf[kGridMax], lif[kGridMax];
    float lr2l[kGridMax], li2l[kGridMax];    // float32 residuals lr2 - lr2f, li2 - li2f (compensated exponent offset)
    int ca[kGridCorrMax], cb[kGridCorrMax];
    float cw[kGridCorrMax];                  // multiplicity - 1 (0: unused slot)
    float dpos_r[kGridMax], dneg_r[kGridMax], dpos_i[kGridMax], dneg_i[kGridMax];   // (level - top/bottom level) log2 e
    float cmag[4];                           // |corner level|: re < 0, re >= 0, im < 0, im >= 0
    int corner_ok, corner_k[4];              // first table index of the corner (re < 0 ? min : max, im < 0 ? min : max), slot 2 (re < 0) + (im < 0)
};

struct BampArgs {
    Geom g;
    DevAlphabet al;
    DevGrid grid;
    const float2* H;
    long long H_stride;          // complex elements between frames, 0 = shared
    const float2* y;
    float sigma2;
    const float* sigma2_pf;
    LossIO io;
    float2* xmap;
    float2* xmmse;
    float* var;
    int* iters;
    float* traj;
    long long frames;
    int stage_H;
    // structured (block-Toeplitz) operator of ampsm_bamp_detect_taps, generic kernel only; taps == nullptr: dense H
    const float2* taps;          // [Lh][Nr][Nt], block (i, j) of H = taps[i - j]
    long long taps_stride;       // complex elements between frames, 0 = shared
    int Lh;
    int cyclic;                  // 1: i - j is taken modulo Lin (channel_truncation 'cyclic', channel.py:67-72)
};

struct VampArgs {
    Geom g;
    DevAlphabet al;
    DevGrid grid;
    const void* U;               // [n][R] complex64 / complex128
    long long U_stride;
    const void* s;               // [R] float / double
    long long s_stride;
    const void* Vh;              // [R][N]
    long long Vh_stride;
    const void* y;               // [frames][n]
    double sigma2_d;             // python float in the reference (vamp.py:19)
    const float* sigma2_pf;
    double sparsity;
    LossIO io;
    void* xmap;                  // complex64, or complex128 in the double variant
    float2* xmmse;
    float* var;
    int* iters;
    float* traj;
    long long frames;
    int stage_Vh;
    float damping;               // vamp2.py only: rho of VAMPLayer (vamp2.py:30, 50)
    unsigned opaque_zero;        // always 0; the kernels use it for scheduling ties the compiler cannot fold (fastops.cuh chain_tie)
};

struct ScampArgs {
    Geom g;
    DevAlphabet al;
    const float* W;              // [Lout][Lin]
    const float2* A;             // [n][N]  (nullptr when the design matrix is given by its taps)
    const float2* taps;          // [Lh][Nr][Nt]: block (r, c) of A = taps[r - c] for 0 <= r - c < Lh (channel.py:89-91), or nullptr
    int Lh;
    const float2* y;             // [frames][n]
    float sigma2;
    const float* sigma2_pf;
    LossIO io;
    float2* xmap;
    float2* xmmse;
    float* psi;
    int* iters;
    float* traj;
    long long frames;
    void* workspace;
};

struct LossArgs {
    Geom g;
    DevAlphabet al;
    const float2* xmap;
    const float2* xmmse;
    const int* iters;
    LossIO io;
    long long frames;
};

struct ShrinkArgs {
    DevAlphabet al;
    float P0, Ps;                // config.P0, config.Ps as float32 tensors (shrink.py:18)
    const float2* r;             // [elems] complex64
    const float* cov;            // [elems] (cov_stride = 1) or one value (cov_stride = 0)
    long long cov_stride;
    long long elems;
    int M;                       // section size (sw_shrinkOOK)
    float2* out_c;
    float* out_f;
    double* sum;
};

// On-device frame generation (framegen.cuh): Philox key / counters, scalings, optional Kronecker roots, the message alphabet
// and the ground-truth outputs the Loss needs.
struct GenArgs {
    unsigned long long seed;     // Philox key
    long long counter_base;      // global number of frame 0 of the call (Philox counter), so shards and chunks draw disjoint frames
    float h_std;                 // std of Re / Im of an i.i.d. channel entry: sqrt(1 / Nr / 2) (channel.py:55)
    float noise_std;             // sqrt(sigma^2 / 2) (channel.py:113-115)
    const float2* Rr_root;       // [n][n] or nullptr;  H = Rr_root G Rt_root (BASELINE config 5)
    const float2* Rt_root;       // [N][N] or nullptr
    int real_roots;              // both roots have zero imaginary parts: half the products
    float ar_t, ar_t_c, ar_r, ar_r_c;   // exponential correlation as AR(1) recursions: rho and sqrt(1 - rho^2), transmit / receive side (0: off)
    int K;
    float2 sym[AMPSM_MAX_K];     // config.symbols as complex64
    long long gray[AMPSM_MAX_K];
    float2* x_out;               // [frames][N] the transmitted vectors (data.py:88)
    long long* idx_out;          // [frames][L] flat non-zero positions (data.py:90)
    long long* sym_out;          // [frames][L] Gray labels (data.py:89)
};

int launch_bamp_generic(const BampArgs& a, bool exp64, cudaStream_t stream);
int launch_bamp_fast(const BampArgs& a, cudaStream_t stream);       // AMPSM_ENOFIT when the shape has no fast path
int launch_vamp_generic(const VampArgs& a, bool is_double, bool exp64, cudaStream_t stream);
int launch_vamp_fast(const VampArgs& a, cudaStream_t stream);       // complex64 32 x 64 factors, else AMPSM_ENOFIT
int launch_vamp_quad(const VampArgs& a, cudaStream_t stream);
int launch_vamp2(const VampArgs& a, bool exp64, cudaStream_t stream);   // vamp2.py, the damped direct form (generic kernel only)
int launch_vamp_dbl(const VampArgs& a, cudaStream_t stream);        // complex128 64 x 128 factors in registers (FP64 pipe), else AMPSM_ENOFIT
int launch_scamp(const ScampArgs& a, bool exp64, cudaStream_t stream);
long long scamp_workspace_bytes(const Geom& g, long long frames);
long long scamp_taps_workspace_bytes(const Geom& g, long long frames, int Lh);   // AMPSM_ENOFIT when the shape has no structured path
int launch_loss(const LossArgs& a, cudaStream_t stream);
int launch_shrink(const ShrinkArgs& a, int kind, cudaStream_t stream);   // kind 0 bayes, 1 shrinkOOK, 2 sw_shrinkOOK
// batched thin SVD H = U diag(s) Vh of dense [frames][n][N] complex64 matrices, n <= 32 (svd_jacobi.cu); sweeps optional
int launch_svd_jacobi(const float2* H, long long frames, int n, int N, float2* U, float* S, float2* Vh, int* sweeps, const float2* y,
                      float2* yrot, cudaStream_t stream);      // y != nullptr: yrot = U^H y instead of U
int launch_identity(float2* I, int n, cudaStream_t stream);
// gen != nullptr: the frames are generated inside the kernel (H == y == nullptr): s, Vh, yrot and gen's ground truth come out
int launch_svd_jacobi_gen(const GenArgs& gen, const Geom& g, long long frames, float* S, float2* Vh, float2* yrot, cudaStream_t stream);
// the same frames written out: H [frames][n][N], y [frames][n] (either may be nullptr) and gen's ground truth
int launch_generate_frames(const GenArgs& gen, const Geom& g, long long frames, float2* H, float2* y, cudaStream_t stream);
int probe_fp32(int device, double* tflops);
int probe_fp32x2(int device, double* tflops);
int probe_fp64(int device, double* tflops);

}  // namespace ampsm
