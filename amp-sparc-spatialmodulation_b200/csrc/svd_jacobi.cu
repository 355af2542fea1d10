// Batched thin SVD of wide complex64 matrices by one-sided (Hestenes) Jacobi: ONE WARP PER MATRIX, matrix in shared memory.
//
// Replaces the caller-side `torch.linalg.svd(A, full_matrices=False)` of the reference's VAMP driver
// (vamp_model.py:56-58) for per-frame channel matrices: H [n][N] (n <= 32 rows, n <= N) -> U [n][n], s [n] (descending),
// Vh [n][N] with H = U diag(s) Vh.  VAMP only ever uses V f(S) V^H and V S U^H (vamp.py:22,67,72), so the phases of the
// singular-vector pairs are free; they differ from LAPACK's.
//
// Method: the rows of W = H are orthogonalised in place by plane rotations (no Gram matrix: the condition number is
// not squared, float32 is enough also for correlated channels); the same rotations applied to Q = I give Q = U^H, so
// that W = U^H H = diag(s) Vh at convergence.  A sweep visits all n(n-1)/2 row pairs in the round-robin (tournament)
// order: 31 steps of 16 disjoint pairs.  Two lanes serve a pair -- lane (j,h) takes the columns c = h (mod 2) of both
// rows: it loads them once into registers (conflict-free 8-byte accesses, padded rows), forms the three inner products,
// combines them with its partner by one shuffle round, rotates in registers and stores back.  A pair whose rows are
// already orthogonal to 4e-7 (relative) is skipped; the sweep loop ends when a whole sweep rotated nothing.
#include "kernels.h"

namespace ampsm {

namespace {

constexpr int kSvdRows = 32;              // positions of the tournament (rows are zero-padded up to it)
constexpr int kSvdWarps = 3;              // warps (matrices) per CTA
constexpr int kSvdMaxSweeps = 16;
constexpr float kSvdTol = 4.0e-7f;

template <int NC>
struct SvdShape {
    static constexpr int wstride = NC + 1;                    // complex elements per row of W (odd: conflict-free columns)
    static constexpr int qstride = kSvdRows + 1;
    static constexpr int warp_bytes = (kSvdRows * wstride + kSvdRows * qstride) * 8 + 2 * kSvdRows * 4;   // + perm, 1/s
};
// column k of a lane's share: lane half h takes the columns whose index mod 16 lies in [8h, 8h+8), so that the 16 lanes
// of a half-warp (8 consecutive rows x 2 halves, row stride = 1 mod 16 in 8-byte units) hit 16 distinct bank pairs
template <int NC>
__device__ __forceinline__ constexpr int svd_col(int c, int h) {
    return NC >= 16 ? ((c >> 3) * 16 + 8 * h + (c & 7)) : (2 * c + h);
}

// NC = number of columns (compile time), n = number of rows (run time, <= 32)
template <int NC>
__global__ void __launch_bounds__(kSvdWarps * 32) svd_jacobi_kernel(const float2* __restrict__ H, long long frames, int n, float2* __restrict__ U,
                                                                   float* __restrict__ S, float2* __restrict__ Vh, int* __restrict__ sweeps_out) {
    using Sh = SvdShape<NC>;
    constexpr int HC = NC / 2, HQ = kSvdRows / 2;             // columns of W / Q per lane
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    float2* W = reinterpret_cast<float2*>(smem + (size_t)wic * Sh::warp_bytes);
    float2* Q = W + kSvdRows * Sh::wstride;
    int* perm = reinterpret_cast<int*>(Q + kSvdRows * Sh::qstride);
    float* sinv = reinterpret_cast<float*>(perm + kSvdRows);
    const int j = lane >> 1, h = lane & 1;

    for (long long f = (long long)blockIdx.x * kSvdWarps + wic; f < frames; f += (long long)gridDim.x * kSvdWarps) {
        // ---- load: W = H (rows >= n are zero), Q = I
        const float2* Hf = H + f * (long long)n * NC;
        for (int e = lane; e < kSvdRows * NC; e += 32) {
            const int r = e / NC, c = e - r * NC;
            W[r * Sh::wstride + c] = r < n ? __ldg(Hf + e) : make_float2(0.f, 0.f);
        }
        for (int e = lane; e < kSvdRows * kSvdRows; e += 32) {
            const int r = e >> 5, c = e & 31;
            Q[r * Sh::qstride + c] = make_float2(r == c ? 1.f : 0.f, 0.f);
        }
        __syncwarp();

        int sweeps = 0;
        for (; sweeps < kSvdMaxSweeps; ++sweeps) {
            bool rotated = false;
            for (int step = 0; step < kSvdRows - 1; ++step) {
                // tournament pairing: position 31 stays, the others rotate
                int p, q;
                if (j == 0) {
                    p = kSvdRows - 1;
                    q = step;
                } else {
                    p = step + j;
                    if (p >= kSvdRows - 1) p -= kSvdRows - 1;
                    q = step - j;
                    if (q < 0) q += kSvdRows - 1;
                }
                float2* wa = W + p * Sh::wstride;
                float2* wb = W + q * Sh::wstride;
                float2 a[HC], b[HC];
                float alpha = 0.f, beta = 0.f, gr = 0.f, gi = 0.f;
#pragma unroll
                for (int c = 0; c < HC; ++c) {
                    a[c] = wa[svd_col<NC>(c, h)];
                    b[c] = wb[svd_col<NC>(c, h)];
                    alpha = fmaf(a[c].x, a[c].x, fmaf(a[c].y, a[c].y, alpha));
                    beta = fmaf(b[c].x, b[c].x, fmaf(b[c].y, b[c].y, beta));
                    gr = fmaf(a[c].x, b[c].x, fmaf(a[c].y, b[c].y, gr));          // gamma = sum conj(a) b
                    gi = fmaf(a[c].x, b[c].y, fmaf(-a[c].y, b[c].x, gi));
                }
                alpha += __shfl_xor_sync(0xffffffffu, alpha, 1);
                beta += __shfl_xor_sync(0xffffffffu, beta, 1);
                gr += __shfl_xor_sync(0xffffffffu, gr, 1);
                gi += __shfl_xor_sync(0xffffffffu, gi, 1);
                const float g2 = gr * gr + gi * gi;
                const bool rot = g2 > kSvdTol * kSvdTol * alpha * beta && g2 > 0.f;
                if (rot) {
                    // b~ = e^{-i phi} b makes <a, b~> = |gamma| real; then the real Jacobi rotation:
                    // zeta = (beta - alpha) / (2 |gamma|), t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), c = 1/sqrt(1+t^2), s = c t
                    const float gabs = sqrtf(g2), ginv = 1.0f / gabs;
                    const float pr = gr * ginv, pi = -gi * ginv;                    // e^{-i phi} = conj(gamma) / |gamma|
                    const float zeta = (beta - alpha) * (0.5f * ginv);
                    const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
                    const float cs = rsqrtf(fmaf(t, t, 1.0f)), sn = cs * t;
                    const float spr = sn * pr, spi = sn * pi, cpr = cs * pr, cpi = cs * pi;
                    // a' = c a - (s p) b ; b' = s a + (c p) b
#pragma unroll
                    for (int c = 0; c < HC; ++c) {
                        const float2 x = a[c], y = b[c];
                        wa[svd_col<NC>(c, h)] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                        wb[svd_col<NC>(c, h)] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                    }
                    float2* qa = Q + p * Sh::qstride;
                    float2* qb = Q + q * Sh::qstride;
#pragma unroll
                    for (int c = 0; c < HQ; ++c) {
                        const int cq = svd_col<kSvdRows>(c, h);
                        const float2 x = qa[cq], y = qb[cq];
                        qa[cq] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                        qb[cq] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                    }
                    rotated = true;
                }
                __syncwarp();
            }
            if (!__any_sync(0xffffffffu, rotated)) {
                ++sweeps;
                break;
            }
        }

        // ---- singular values = row norms of W (lane r owns row r), descending order by rank counting
        float nrm = 0.f;
#pragma unroll 8
        for (int c = 0; c < NC; ++c) {
            const float2 w = W[lane * Sh::wstride + c];
            nrm = fmaf(w.x, w.x, fmaf(w.y, w.y, nrm));
        }
        const float sv = sqrtf(nrm);
        int rank = 0;
#pragma unroll
        for (int o = 0; o < 32; ++o) {
            const float other = __shfl_sync(0xffffffffu, sv, o);
            rank += (other > sv) || (other == sv && o < lane);
        }
        // row `lane` of W / Q becomes singular triplet `rank`; rows beyond n (zero padding) rank last and are dropped
        perm[rank] = lane;
        sinv[rank] = sv > 0.f ? 1.0f / sv : 0.f;
        if (rank < n) S[f * n + rank] = sv;
        __syncwarp();
        // Vh[k][:] = W[row_k][:] / s_k  (lanes sweep the columns), U[i][k] = conj(Q[row_k][i])  (lanes sweep k): coalesced stores
        for (int k = 0; k < n; ++k) {
            const int row = perm[k];
            const float sc = sinv[k];
            float2* out = Vh + (f * n + k) * (long long)NC;
            for (int c = lane; c < NC; c += 32) {
                const float2 w = W[row * Sh::wstride + c];
                out[c] = make_float2(w.x * sc, w.y * sc);
            }
        }
        if (lane < n) {
            const int row = perm[lane];
            for (int i = 0; i < n; ++i) {
                const float2 qv = Q[row * Sh::qstride + i];
                U[(f * n + i) * (long long)n + lane] = make_float2(qv.x, -qv.y);
            }
        }
        if (sweeps_out && lane == 0) sweeps_out[f] = sweeps;
        __syncwarp();
    }
}

template <int NC>
int launch_svd_nc(const float2* H, long long frames, int n, float2* U, float* S, float2* Vh, int* sweeps, cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = svd_jacobi_kernel<NC>;
    const size_t smem = (size_t)SvdShape<NC>::warp_bytes * kSvdWarps;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(svd)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSvdWarps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (frames + kSvdWarps - 1) / kSvdWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kSvdWarps * 32, smem, stream>>>(H, frames, n, U, S, Vh, sweeps);
    count_launch();
    return check_cuda(cudaGetLastError(), "svd_jacobi_kernel launch");
}

}  // namespace

int launch_svd_jacobi(const float2* H, long long frames, int n, int N, float2* U, float* S, float2* Vh, int* sweeps, cudaStream_t stream) {
    if (n < 1 || n > kSvdRows || N < n) {
        set_error("batched SVD: needs 1 <= n <= 32 rows and n <= N columns (got %d x %d)", n, N);
        return AMPSM_ENOFIT;
    }
    switch (N) {
        case 8: return launch_svd_nc<8>(H, frames, n, U, S, Vh, sweeps, stream);
        case 16: return launch_svd_nc<16>(H, frames, n, U, S, Vh, sweeps, stream);
        case 32: return launch_svd_nc<32>(H, frames, n, U, S, Vh, sweeps, stream);
        case 64: return launch_svd_nc<64>(H, frames, n, U, S, Vh, sweeps, stream);
        default:
            set_error("batched SVD: column counts 8, 16, 32, 64 are instantiated (got %d)", N);
            return AMPSM_ENOFIT;
    }
}

}  // namespace ampsm
