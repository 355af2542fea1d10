// SCAMP for a batch of frames that share one design matrix A (scamp.py:8-25, 43-59, 61-68, 77-108).
// With A shared, the two mat-vecs of every frame become complex GEMMs over the frame batch
//   residual : Z  = Y - Xh A^T + rep(b) o Z          (scamp.py:48)
//   estimate : Xm = Xh + rep(tau) o ((Z / phi) conj(A)) (scamp.py:56)
// A is streamed tile by tile; tiles of A that are entirely zero (the band structure of a coupled design matrix,
// channel.py:89-91) are skipped through a tile map computed once per call from A itself.  Frames that met the
// exit test are frozen: their tiles are skipped and their state is never rewritten.
#include <cstdlib>

#include "blockops.cuh"
#include "kernels.h"
#include "scamp_ws.cuh"

namespace ampsm {

constexpr int BM = 64;   // frames per tile
constexpr int BN = 32;   // output columns per tile (rows of A in `residual`, columns of A in `estimate`)
constexpr int BK = 32;   // reduction chunk

__host__ __device__ inline size_t up256(size_t v) { return (v + 255) & ~size_t(255); }

static bool scamp_use_tc(const Geom& g, long long F) { return F >= 128 && !getenv("AMPSM_SCAMP_SIMT") && scamp_tc_fits(g.n, g.N, F); }

// Lh > 0: structured path (design matrix given by its Lh taps, scamp_st.cu): Zs padded per frame, pre-split tap planes
static size_t ws_layout(const Geom& g, long long F, bool exp64, bool own_xmap, ScampWs* ws, unsigned char* base, int Lh = 0) {
    const ScampStPlan sp = Lh > 0 ? scamp_st_plan(g, Lh) : ScampStPlan{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t at = o;
        o = up256(o + bytes);
        return base ? base + at : nullptr;
    };
    ScampWs w{};
    w.Xh = (float2*)take((size_t)F * g.N * 8);
    w.Z = (float2*)take((size_t)F * g.n * 8);
    w.Zs = (float2*)take((size_t)F * (Lh > 0 ? sp.zs_stride : g.n) * 8);
    w.Xmap = (float2*)take(own_xmap ? (size_t)F * g.N * 8 : 0);
    w.psi = (float*)take((size_t)F * g.Lin * 4);
    w.phi = (float*)take((size_t)F * g.Lout * 4);
    w.tau = (float*)take((size_t)F * g.Lin * 4);
    w.b = (float*)take((size_t)F * g.Lout * 4);
    w.active = (int*)take((size_t)F * 4);
    w.iters = (int*)take((size_t)F * 4);
    w.scr = take((size_t)F * g.N * 3 * (exp64 ? 8 : 4));
    w.nzc = (g.N + TILE - 1) / TILE;
    w.nz = take(Lh > 0 ? 0 : (size_t)((g.n + TILE - 1) / TILE) * w.nzc);
    w.At = (float2*)take(Lh == 0 && scamp_use_tc(g, F) ? (size_t)g.n * g.N * 8 : 0);     // A^T for the tensor-core `estimate` GEMM
    w.bplanes0 = take(Lh > 0 ? sp.bplane_bytes0 : 0);
    w.bplanes1 = take(Lh > 0 ? sp.bplane_bytes1 : 0);
    w.notclose = (int*)take((size_t)F * 4);
    if (ws) *ws = w;
    return o;
}

long long scamp_workspace_bytes(const Geom& g, long long frames) { return (long long)ws_layout(g, frames, true, true, nullptr, nullptr); }
long long scamp_taps_workspace_bytes(const Geom& g, long long frames, int Lh) {
    if (!scamp_st_plan(g, Lh).ok) return AMPSM_ENOFIT;
    return (long long)ws_layout(g, frames, true, true, nullptr, nullptr, Lh);
}

// ---- zero-tile map of A -------------------------------------------------------------------------------------
__global__ void scamp_nzmap_kernel(const float2* __restrict__ A, int n, int N, unsigned char* nz, int nzc) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    bool any = false;
    for (int e = threadIdx.x; e < TILE * TILE; e += blockDim.x) {
        const int i = ti * TILE + e / TILE, j = tj * TILE + e % TILE;
        if (i < n && j < N) {
            const float2 v = A[(size_t)i * N + j];
            any |= (v.x != 0.f) || (v.y != 0.f);   // NaN counts as non-zero
        }
    }
    const int r = __syncthreads_or(any ? 1 : 0);
    if (threadIdx.x == 0) nz[ti * nzc + tj] = (unsigned char)(r != 0);
}

// ---- state init (scamp.py:17-25) ---------------------------------------------------------------------------
__global__ void scamp_init_kernel(ScampWs w, Geom g, const float2* __restrict__ y, long long F) {
    const long long f = blockIdx.x;
    if (f >= F) return;
    for (int i = threadIdx.x; i < g.n; i += blockDim.x) w.Z[f * g.n + i] = y[f * g.n + i];
    for (int j = threadIdx.x; j < g.N; j += blockDim.x) w.Xh[f * g.N + j] = make_float2(0.f, 0.f);
    for (int c = threadIdx.x; c < g.Lin; c += blockDim.x) w.psi[f * g.Lin + c] = 1.0f;
    for (int r = threadIdx.x; r < g.Lout; r += blockDim.x) w.phi[f * g.Lout + r] = INFINITY;
    if (threadIdx.x == 0) {
        w.active[f] = 1;
        w.iters[f] = 0;
        w.notclose[f] = 0;
    }
}

// ---- per-frame block scalars: gamma, b, phi, tau (scamp.py:45-52) -------------------------------------------
__global__ void scamp_scalars_kernel(ScampWs w, Geom g, const float* __restrict__ W, float sigma2,
                                     const float* __restrict__ sigma2_pf, long long F) {
    const long long f = blockIdx.x;
    if (f >= F || !w.active[f]) return;
    const int Lr = g.Lout, Lc = g.Lin;
    const float s2 = sigma2_pf ? sigma2_pf[f] : sigma2;
    float* phi = w.phi + f * Lr;
    for (int r = threadIdx.x; r < Lr; r += blockDim.x) {
        float acc = 0.f;
        for (int c = 0; c < Lc; ++c) acc = fmaf(W[r * Lc + c], w.psi[f * Lc + c], acc);
        const float gma = acc / (float)Lc;
        w.b[f * Lr + r] = gma / phi[r];          // old phi (inf on the first pass -> 0)
        phi[r] = s2 + gma;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < Lc; c += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < Lr; ++r) acc = fmaf(W[r * Lc + c], __frcp_rn(phi[r]), acc);
        w.tau[f * Lc + c] = (float)g.L / acc / (float)g.Nr;      // L = Na*Lin, Mr = Nr (scamp.py:52)
    }
}

// ---- batched complex GEMM with fused epilogues ------------------------------------------------------------------
// MODE 0 (residual): C[f][i] = sum_j Xh[f][j] A[i][j];      MODE 1 (estimate): C[f][j] = sum_i Zs[f][i] conj(A[i][j])
template <int MODE>
__global__ void __launch_bounds__(256) scamp_gemm_kernel(ScampWs w, Geom g, const float2* __restrict__ A,
                                                         const float2* __restrict__ y, long long F) {
    __shared__ float2 Xs[BM][BK + 1];
    __shared__ float2 As[BN][BK + 1];      // MODE 0: [out col][k];  MODE 1: [k][out col] (BN == BK)
    __shared__ int act_s[BM];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long f0 = (long long)blockIdx.y * BM;
    const int c0 = blockIdx.x * BN;                       // first output column of this tile
    const int Kdim = MODE == 0 ? g.N : g.n;               // reduction length
    const int Odim = MODE == 0 ? g.n : g.N;               // output width
    const float2* __restrict__ Xin = MODE == 0 ? w.Xh : w.Zs;

    int any_active = 0;
    if (tid < BM) {
        const long long f = f0 + tid;
        const int a = (f < F) ? w.active[f] : 0;
        act_s[tid] = a;
        any_active = a;
    }
    if (!__syncthreads_or(any_active)) return;

    float2 acc[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = make_float2(0.f, 0.f);

    for (int k0 = 0; k0 < Kdim; k0 += BK) {
        const int trow = MODE == 0 ? c0 / TILE : k0 / TILE;
        const int tcol = MODE == 0 ? k0 / TILE : c0 / TILE;
        if (!w.nz[trow * w.nzc + tcol]) continue;       // uniform across the block
        // stage the frame tile (BM x BK) and the A tile (BN x BK or BK x BN), zero-filled at the edges
        for (int e = tid; e < BM * BK; e += 256) {
            const int r = e / BK, c = e % BK;
            const long long f = f0 + r;
            const int k = k0 + c;
            Xs[r][c] = (f < F && k < Kdim) ? Xin[f * Kdim + k] : make_float2(0.f, 0.f);
        }
        for (int e = tid; e < BN * BK; e += 256) {
            const int r = e / BK, c = e % BK;
            int i, j;
            if (MODE == 0) { i = c0 + r; j = k0 + c; } else { i = k0 + r; j = c0 + c; }
            As[r][c] = (i < g.n && j < g.N) ? A[(size_t)i * g.N + j] : make_float2(0.f, 0.f);
        }
        __syncthreads();
#pragma unroll 8
        for (int k = 0; k < BK; ++k) {
            float2 a0, a1;
            if (MODE == 0) { a0 = As[tx][k]; a1 = As[tx + 16][k]; } else { a0 = As[k][tx]; a1 = As[k][tx + 16]; }
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 x = Xs[ty * 4 + r][k];
                if (MODE == 0) {
                    acc[r][0].x = fmaf(x.x, a0.x, fmaf(-x.y, a0.y, acc[r][0].x));
                    acc[r][0].y = fmaf(x.x, a0.y, fmaf(x.y, a0.x, acc[r][0].y));
                    acc[r][1].x = fmaf(x.x, a1.x, fmaf(-x.y, a1.y, acc[r][1].x));
                    acc[r][1].y = fmaf(x.x, a1.y, fmaf(x.y, a1.x, acc[r][1].y));
                } else {   // x * conj(a)
                    acc[r][0].x = fmaf(x.x, a0.x, fmaf(x.y, a0.y, acc[r][0].x));
                    acc[r][0].y = fmaf(x.y, a0.x, fmaf(-x.x, a0.y, acc[r][0].y));
                    acc[r][1].x = fmaf(x.x, a1.x, fmaf(x.y, a1.y, acc[r][1].x));
                    acc[r][1].y = fmaf(x.y, a1.x, fmaf(-x.x, a1.y, acc[r][1].y));
                }
            }
        }
        __syncthreads();
    }
    // epilogue
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int fr = ty * 4 + r;
        const long long f = f0 + fr;
        if (f >= F || !act_s[fr]) continue;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int o = c0 + tx + 16 * c;
            if (o >= Odim) continue;
            const float2 s = acc[r][c];
            if (MODE == 0) {
                const int blk = o / g.Nr;                                  // row block (Mr = Nr)
                const float2 yv = y[f * g.n + o], zo = w.Z[f * g.n + o];
                const float bb = w.b[f * g.Lout + blk];
                const float2 zn = make_float2(yv.x - s.x + bb * zo.x, yv.y - s.y + bb * zo.y);   // scamp.py:48
                w.Z[f * g.n + o] = zn;
                w.Zs[f * g.n + o] = cdiv_real(zn, w.phi[f * g.Lout + blk]);                      // z / phi_use
            } else {
                const float tau = w.tau[f * g.Lin + o / g.Nt];                                   // column block (Mc = Nt)
                const float2 xo = w.Xh[f * g.N + o];
                w.Xmap[f * g.N + o] = make_float2(fmaf(tau, s.x, xo.x), fmaf(tau, s.y, xo.y));   // scamp.py:56
            }
        }
    }
}

// ---- denoiser (mean only), psi update, exit test (scamp.py:57-59, 105) ------------------------------------------
template <bool EXP64>
__global__ void __launch_bounds__(256) scamp_denoise_kernel(ScampWs w, Geom g, DevAlphabet al, const float2* x_true,
                                                            float* traj, int t, long long F) {
    using E = typename ExpT<EXP64>::type;
    __shared__ double red[96];
    __shared__ int close_s;
    const long long f = blockIdx.x;
    if (f >= F || !w.active[f]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float2* xmap = w.Xmap + f * g.N;
    float2* xh = w.Xh + f * g.N;
    const float* tau = w.tau + f * g.Lin;
    E* scr = reinterpret_cast<E*>(w.scr) + (size_t)f * 3 * g.N;
    double gshift = 0.0;
    if (EXP64 && g.shift_mode == 1) gshift = block_absmax_exponent<float2>(g, al, xmap, tau, 0.f, true, red, g.Nt);
    if (tid == 0) close_s = 1;
    block_denoise<EXP64, float2>(g, al, xmap, tau, 0.f, true, gshift, xh, nullptr, scr, g.Nt);
    __syncthreads();
    // psi_c = 1 - sum_{block c} |xh|^2 / Na ; exit when allclose(psi_new, psi_old)
    double psi_sum = 0.0;
    for (int c = warp; c < g.Lin; c += nwarps) {
        float acc = 0.f;
        for (int j = lane; j < g.Nt; j += 32) {
            const float2 v = xh[c * g.Nt + j];
            acc += v.x * v.x + v.y * v.y;
        }
        acc = warp_sum(acc);
        if (lane == 0) {
            const float pn = 1.0f - acc / (float)g.Na;
            const float po = w.psi[f * g.Lin + c];
            if (!(fabsf(pn - po) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, po))))) atomicAnd(&close_s, 0);
            w.psi[f * g.Lin + c] = pn;
            psi_sum += pn;
        }
    }
    if (traj) {
        double tau_sum = 0.0, mse = 0.0;
        for (int c = tid; c < g.Lin; c += blockDim.x) tau_sum += tau[c];
        if (x_true)
            for (int j = tid; j < g.N; j += blockDim.x) {
                const float2 xe = xh[j], xt = x_true[f * g.N + j];
                const double dr = (double)xe.x - xt.x, di = (double)xe.y - xt.y;
                mse += dr * dr + di * di;
            }
        tau_sum = warp_sum(tau_sum);
        mse = warp_sum(mse);
        psi_sum = warp_sum(psi_sum);
        __syncthreads();
        if (lane == 0) {
            red[warp] = tau_sum;
            red[32 + warp] = psi_sum;
            red[64 + warp] = mse;
        }
        __syncthreads();
        if (tid == 0) {
            double a = 0, b = 0, c = 0;
            for (int i = 0; i < nwarps; ++i) {
                a += red[i];
                b += red[32 + i];
                c += red[64 + i];
            }
            float* tr = traj + (f * g.max_iters + t) * 3;
            tr[0] = (float)(a / g.Lin);
            tr[1] = (float)(b / g.Lin);
            tr[2] = (float)(c / g.N);
        }
    }
    __syncthreads();
    if (tid == 0) {
        w.iters[f] = t + 1;
        if (g.early_exit && close_s) w.active[f] = 0;
    }
}

// ---- fast denoiser (float32 exp, per-section shift, no trajectory): one CTA per (column block, frame), one warp per section.
// Mean only (scamp.py:61-68): exponents as float64 products, float32 ex2 of the shifted difference, everything in registers --
// no scratch round trip through global memory (the generic kernel parks 3 N values per frame there).  psi of the block
// (scamp.py:59) and its allclose test (scamp.py:105) are fused; a frame is retired by scamp_exit_kernel once none of its
// blocks raised `notclose`.
template <int MA>     // antennas per lane, M <= 32 MA
__global__ void __launch_bounds__(256) scamp_denoise_fast_kernel(ScampWs w, Geom g, DevAlphabet al, long long F) {
    __shared__ float wsum[8];
    const long long f = blockIdx.y;
    const int c = blockIdx.x;                                  // column block (Mc = Nt columns = Na sections)
    if (f >= F || !w.active[f]) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int M = g.M, K = al.K;
    const float tau = w.tau[f * g.Lin + c];
    const float rt = __frcp_rn(tau / 2.0f);                    // s / (tau / 2) as complex64 / real (scamp.py:63)
    const float2* xmap = w.Xmap + f * g.N + (size_t)c * g.Nt;
    float2* xh = w.Xh + f * g.N + (size_t)c * g.Nt;
    float energy = 0.f;
    for (int sec = warp; sec < g.Na; sec += 8) {
        double qr[MA], qi[MA];
        float lmax = -INFINITY;
#pragma unroll
        for (int a = 0; a < MA; ++a) {
            const int m = lane + 32 * a;
            qr[a] = qi[a] = 0.0;
            if (m < M) {
                const float2 s = xmap[sec * M + m];
                const float q_r = __fmul_rn(s.x, rt), q_i = __fmul_rn(s.y, rt);
                qr[a] = (double)q_r;
                qi[a] = (double)q_i;
                for (int k = 0; k < K; ++k) lmax = fmaxf(lmax, fmaf(q_r, al.ref[k], q_i * al.imf[k]));
            }
        }
        const double shift = (double)warp_max(lmax);           // a common shift of the section, nothing else
        float s0 = 0.f, s1r[MA], s1i[MA];
#pragma unroll
        for (int a = 0; a < MA; ++a) {
            s1r[a] = s1i[a] = 0.f;
            if (lane + 32 * a < M) {
                for (int k = 0; k < K; ++k) {
                    const double x = fma(qr[a], al.re[k], qi[a] * al.im[k]);
                    const float e = exp2f((float)(x - shift) * 1.4426950408889634f);
                    s0 += e;
                    s1r[a] = fmaf(al.ref[k], e, s1r[a]);
                    s1i[a] = fmaf(al.imf[k], e, s1i[a]);
                }
            }
        }
        const float rz = 1.0f / warp_sum(s0);
#pragma unroll
        for (int a = 0; a < MA; ++a) {
            const int m = lane + 32 * a;
            if (m < M) {
                const float2 v = make_float2(s1r[a] * rz, s1i[a] * rz);
                xh[sec * M + m] = v;
                energy = fmaf(v.x, v.x, fmaf(v.y, v.y, energy));
            }
        }
    }
    energy = warp_sum(energy);
    if (lane == 0) wsum[warp] = energy;
    __syncthreads();
    if (tid == 0) {
        float acc = 0.f;
        for (int i = 0; i < 8; ++i) acc += wsum[i];
        const float pn = 1.0f - acc / (float)g.Na;             // scamp.py:59
        const float po = w.psi[f * g.Lin + c];
        if (!(fabsf(pn - po) <= __fadd_rn(kAtol, fabsf(__fmul_rn(kRtol, po))))) w.notclose[f] = 1;
        w.psi[f * g.Lin + c] = pn;
    }
}

__global__ void scamp_exit_kernel(ScampWs w, Geom g, int t, long long F) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F || !w.active[f]) return;
    w.iters[f] = t + 1;
    if (g.early_exit && !w.notclose[f]) w.active[f] = 0;       // scamp.py:105
    w.notclose[f] = 0;
}

__global__ void scamp_finish_kernel(ScampWs w, Geom g, float2* xmmse, float* psi, int* iters, float* traj, long long F) {
    const long long f = blockIdx.x;
    if (f >= F) return;
    if (xmmse)
        for (int j = threadIdx.x; j < g.N; j += blockDim.x) xmmse[f * g.N + j] = w.Xh[f * g.N + j];
    if (psi)
        for (int c = threadIdx.x; c < g.Lin; c += blockDim.x) psi[f * g.Lin + c] = w.psi[f * g.Lin + c];
    const int done = w.iters[f];
    if (iters && threadIdx.x == 0) iters[f] = done;
    if (traj && done > 0)
        for (int t = done + threadIdx.x; t < g.max_iters; t += blockDim.x)
            for (int q = 0; q < 3; ++q) traj[(f * g.max_iters + t) * 3 + q] = traj[(f * g.max_iters + done - 1) * 3 + q];
}

int launch_scamp(const ScampArgs& a, bool exp64, cudaStream_t stream) {
    const Geom& g = a.g;
    const long long F = a.frames;
    if (F <= 0) return 0;
    const bool own_xmap = a.xmap == nullptr;
    const int Lh = a.taps ? a.Lh : 0;
    ScampStPlan sp{};
    if (Lh > 0) {
        sp = scamp_st_plan(g, Lh);
        if (!sp.ok) { set_error("SCAMP taps: shape does not fit the structured tensor-core path (Lin <= 128, Lh Nr <= 128, Nr even)"); return AMPSM_ENOFIT; }
    }
    unsigned char* base = reinterpret_cast<unsigned char*>(a.workspace);
    bool own_ws = false;
    if (!base) {
        const size_t bytes = ws_layout(g, F, exp64, own_xmap, nullptr, nullptr, Lh);
        if (int e = check_cuda(cudaMallocAsync((void**)&base, bytes, stream), "cudaMallocAsync(scamp workspace)")) return e;
        own_ws = true;
    }
    ScampWs w;
    ws_layout(g, F, exp64, own_xmap, &w, base, Lh);
    if (!own_xmap) w.Xmap = a.xmap;
    const int tr = (g.n + TILE - 1) / TILE, tc = (g.N + TILE - 1) / TILE;
    if (Lh > 0) {
        cudaMemsetAsync(w.Zs, 0, (size_t)F * sp.zs_stride * 8, stream);          // the padding blocks of Zs stay zero
    } else {
        scamp_nzmap_kernel<<<dim3(tc, tr), 128, 0, stream>>>(a.A, g.n, g.N, w.nz, w.nzc);
        count_launch();
    }
    scamp_init_kernel<<<(unsigned)F, 128, 0, stream>>>(w, g, a.y, F);
    count_launch();
    const dim3 grid_res((g.n + BN - 1) / BN, (unsigned)((F + BM - 1) / BM));
    const dim3 grid_est((g.N + BN - 1) / BN, (unsigned)((F + BM - 1) / BM));
    // batches of >= 128 frames: both GEMMs on the tensor cores (tcgen05, 3xTF32, scamp_tc.cu); smaller ones on the SIMT tiles
    const bool use_tc = Lh == 0 && scamp_use_tc(g, F);
    auto fail = [&](int e) {                               // every exit path returns the workspace it allocated
        if (own_ws) cudaFreeAsync(base, stream);
        return e;
    };
    if (use_tc)
        if (int e = scamp_tc_prepare(a.A, w.At, g.n, g.N, stream)) return fail(e);
    if (Lh > 0)
        if (int e = scamp_st_prepare(g, sp, Lh, a.taps, w.bplanes0, w.bplanes1, stream)) return fail(e);
    // float32-exp mode without trajectory: register-resident denoiser (one warp per section) + exit kernel; on the structured
    // path it is fused into the estimate GEMM's epilogue when the section size allows
    const bool fast_dn = !exp64 && g.shift_mode == 0 && !a.traj && g.M <= 128 && g.Nt == g.Na * g.M && F <= 65535 && !getenv("AMPSM_SCAMP_GENERIC_DENOISER");
    const bool fused = Lh > 0 && fast_dn && scamp_st_can_fuse(g, a.al) && !getenv("AMPSM_SCAMP_UNFUSED");
    for (int t = 0; t < g.max_iters; ++t) {
        if (Lh == 0) {                                         // the structured residual kernel computes the block scalars itself
            scamp_scalars_kernel<<<(unsigned)F, 64, 0, stream>>>(w, g, a.W, a.sigma2, a.sigma2_pf, F);
            count_launch();
        }
        if (Lh > 0) {
            if (int e = scamp_st_gemm(0, w, g, sp, Lh, w.bplanes0, a.y, F, a.al, false, a.W, a.sigma2, a.sigma2_pf, t, stream)) return fail(e);
            if (int e = scamp_st_gemm(1, w, g, sp, Lh, w.bplanes1, a.y, F, a.al, fused, a.W, a.sigma2, a.sigma2_pf, t, stream)) return fail(e);
        } else if (use_tc) {
            if (int e = scamp_tc_gemm(0, w, g, a.A, a.y, F, stream)) return fail(e);
            if (int e = scamp_tc_gemm(1, w, g, w.At, a.y, F, stream)) return fail(e);
        } else {
            scamp_gemm_kernel<0><<<grid_res, 256, 0, stream>>>(w, g, a.A, a.y, F);
            scamp_gemm_kernel<1><<<grid_est, 256, 0, stream>>>(w, g, a.A, a.y, F);
            count_launch();
            count_launch();
        }
        if (fused) continue;                                   // denoiser, psi, exit test and retirement ran in the estimate kernel
        count_launch();
        if (fast_dn) {
            const dim3 gd((unsigned)g.Lin, (unsigned)F);
            if (g.M <= 32) scamp_denoise_fast_kernel<1><<<gd, 256, 0, stream>>>(w, g, a.al, F);
            else if (g.M <= 64) scamp_denoise_fast_kernel<2><<<gd, 256, 0, stream>>>(w, g, a.al, F);
            else scamp_denoise_fast_kernel<4><<<gd, 256, 0, stream>>>(w, g, a.al, F);
            scamp_exit_kernel<<<(unsigned)((F + 255) / 256), 256, 0, stream>>>(w, g, t, F);
            count_launch();
        } else if (exp64)
            scamp_denoise_kernel<true><<<(unsigned)F, 256, 0, stream>>>(w, g, a.al, a.io.x_true, a.traj, t, F);
        else
            scamp_denoise_kernel<false><<<(unsigned)F, 256, 0, stream>>>(w, g, a.al, a.io.x_true, a.traj, t, F);
    }
    scamp_finish_kernel<<<(unsigned)F, 128, 0, stream>>>(w, g, a.xmmse, a.psi, a.iters, a.traj, F);
    count_launch();
    int rc = check_cuda(cudaGetLastError(), "scamp kernels launch");
    if (rc == 0 && (a.io.x_true || a.io.counters)) {
        LossArgs la{};
        la.g = g;
        la.al = a.al;
        la.xmap = w.Xmap;
        la.xmmse = w.Xh;
        la.iters = w.iters;
        la.io = a.io;
        la.frames = F;
        rc = launch_loss(la, stream);
    }
    if (own_ws) cudaFreeAsync(base, stream);
    return rc;
}

}  // namespace ampsm
