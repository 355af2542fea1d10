// On-device generation of Monte-Carlo frames with a counter-based generator (Philox4x32-10), one warp per frame, the
// channel matrix written straight into the shared-memory tile the consumer works on.
//
// Replaces, for throughput sweeps, the reference's per-epoch input generation:
//   channel.py:53-55   H = (N(0,1) + j N(0,1)) sqrt(1 / Nr / 2)          -> stream 0, Box-Muller on Philox words
//   data.py:74-91      one active antenna per section, one symbol each   -> stream 2
//   channel.py:113-115 w = (N(0,1) + j N(0,1)) sqrt(sigma^2 / 2), y = H x + w -> stream 1
// plus BASELINE config 5's Kronecker correlation H = Rr^(1/2) G Rt^(1/2) (no generator of the reference draws such channels):
// for the exponential model by two AR(1) recursions, for arbitrary correlation matrices from their roots.
// The draws are NOT the reference's (numpy MT19937 / torch Philox in another order): parity subsets keep the reference's
// own RNG path (tests/golden), this path is pinned by a numpy restatement of the same counter layout (tests/test_framegen.py)
// and by "generate to HBM, then detect" == "generate inside the SVD kernel" bit for bit.
//
// Counter layout: key = seed (64 bit), counter = (frame lo, frame hi, stream, index) with `frame` the global frame number
// (gen.counter_base + frame in call), so shards, chunks and re-runs of a sub-range draw the same frames.
//   stream 0, index i : entries 2 i, 2 i + 1 of G in row-major order (4 normals)
//   stream 1, index i : entries 2 i, 2 i + 1 of the noise vector
//   stream 2, index s : section s: word 0 -> antenna floor(w M / 2^32), word 1 -> symbol floor(w K / 2^32)
#pragma once
#include "kernels.h"

namespace ampsm {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr unsigned M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const unsigned hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// two standard normals from two 32-bit words (Box-Muller on 24-bit uniforms in (0, 1); largest radius 5.9 sigma)
__device__ __forceinline__ float2 box_muller(unsigned a, unsigned b) {
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    __sincosf(6.28318530717958647692f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

// Frame `fg` (global number) of the stream described by `ga` into the warp's tile W (32 rows of `wstride` complex elements,
// rows >= n zeroed) and yv (32 entries, rows >= n zeroed); ground truth of the frame with call-local index `f` to ga.x_out /
// idx_out / sym_out.  NC columns (compile time), n <= 32 rows, L <= 32 sections.  All 32 lanes call it.
// SWZ: the consumer's swizzled row placement (svd_jacobi.cu, SvdShape::swz): row r starts at r wstride + rho(r).
template <int NC, bool SWZ>
__device__ __forceinline__ void gen_frame(const GenArgs& ga, const Geom& g, long long f, int n, float2* W, int wstride, float2* yv, int lane) {
    auto row = [&](int r) { return SWZ ? r * wstride + ((r & 7) ^ ((r & 8) ? 7 : 0)) : r * wstride; };
    const unsigned long long fg = (unsigned long long)(ga.counter_base + f);
    const uint2 key = make_uint2((unsigned)ga.seed, (unsigned)(ga.seed >> 32));
    const unsigned flo = (unsigned)fg, fhi = (unsigned)(fg >> 32);
    // ---- G: i.i.d. CN(0, 2 h_std^2) entries (channel.py:53-55)
    for (int i = lane; i < 16 * NC; i += 32) {                   // 32 rows x NC columns, two entries per Philox block
        const int e = 2 * i, r = e / NC, c = e - r * NC;
        float2 v0 = make_float2(0.f, 0.f), v1 = v0;
        if (r < n) {
            const uint4 w = philox4x32_10(make_uint4(flo, fhi, 0u, (unsigned)i), key);
            v0 = box_muller(w.x, w.y);
            v1 = box_muller(w.z, w.w);
            v0 = make_float2(v0.x * ga.h_std, v0.y * ga.h_std);
            v1 = make_float2(v1.x * ga.h_std, v1.y * ga.h_std);
        }
        W[row(r) + c] = v0;
        W[row(r) + c + 1] = v1;
    }
    __syncwarp();
    // ---- exponential Kronecker correlation (BASELINE config 5), fast form: R[i][j] = rho^|i-j| is the covariance of the AR(1)
    // sequence x_0 = w_0, x_k = rho x_{k-1} + sqrt(1 - rho^2) w_k, so running that recursion along the columns (transmit side) and
    // then along the rows (receive side) of G gives H = A G B with A A^H = Rr, B^H B = Rt -- the distribution of
    // Rr^(1/2) G Rt^(1/2) (G is unitarily invariant) for 6 instead of ~380 FMAs per entry.
    if (ga.ar_t != 0.f) {                                        // lane = row, along the columns
        const float rho = ga.ar_t, cc = ga.ar_t_c;
        float2 prev = W[row(lane)];
        for (int c = 1; c < NC; ++c) {
            const float2 w = W[row(lane) + c];
            prev = make_float2(fmaf(rho, prev.x, cc * w.x), fmaf(rho, prev.y, cc * w.y));
            W[row(lane) + c] = prev;
        }
        __syncwarp();
    }
    if (ga.ar_r != 0.f) {                                        // lane = column(s), along the rows
        const float rho = ga.ar_r, cc = ga.ar_r_c;
        for (int c = lane; c < NC; c += 32) {
            float2 prev = W[row(0) + c];
            for (int r = 1; r < n; ++r) {
                const float2 w = W[row(r) + c];
                prev = make_float2(fmaf(rho, prev.x, cc * w.x), fmaf(rho, prev.y, cc * w.y));
                W[row(r) + c] = prev;
            }
        }
        __syncwarp();
    }
    // ---- Kronecker correlation with arbitrary roots: H = Rr_root G Rt_root, lane = column(s) c = lane + 32 j
    if (ga.Rt_root) {
        constexpr int CJ = (NC + 31) / 32;
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 16) {                    // T = G Rt_root, sixteen rows at a time (in place)
            float2 acc[CJ][16];
#pragma unroll
            for (int j = 0; j < CJ; ++j)
#pragma unroll
                for (int r = 0; r < 16; ++r) acc[j][r] = make_float2(0.f, 0.f);
#pragma unroll 2
            for (int k = 0; k < NC; ++k) {
                float2 t[CJ];
#pragma unroll
                for (int j = 0; j < CJ; ++j) {
                    const int c = lane + 32 * j;
                    t[j] = c < NC ? __ldg(ga.Rt_root + (size_t)k * NC + c) : make_float2(0.f, 0.f);
                }
                if (ga.real_roots) {                             // real roots (exponential correlation): half the products
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float2 gv = W[row(r0 + r) + k];                        // broadcast
#pragma unroll
                        for (int j = 0; j < CJ; ++j) {
                            acc[j][r].x = fmaf(gv.x, t[j].x, acc[j][r].x);
                            acc[j][r].y = fmaf(gv.y, t[j].x, acc[j][r].y);
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 16; ++r) {
                        const float2 gv = W[row(r0 + r) + k];                        // broadcast
#pragma unroll
                        for (int j = 0; j < CJ; ++j) {
                            acc[j][r].x = fmaf(gv.x, t[j].x, fmaf(-gv.y, t[j].y, acc[j][r].x));
                            acc[j][r].y = fmaf(gv.x, t[j].y, fmaf(gv.y, t[j].x, acc[j][r].y));
                        }
                    }
                }
            }
            __syncwarp();                                        // every lane has read rows r0 .. r0 + 15 of G
#pragma unroll
            for (int j = 0; j < CJ; ++j) {
                const int c = lane + 32 * j;
                if (c < NC)
#pragma unroll
                    for (int r = 0; r < 16; ++r) W[row(r0 + r) + c] = acc[j][r];
            }
        }
        __syncwarp();
    }
    if (ga.Rr_root) {                                            // H = Rr_root T: a lane's columns are its own
        constexpr int CJ = (NC + 31) / 32;
#pragma unroll 1
        for (int j = 0; j < CJ; ++j) {
            const int c = lane + 32 * j;
            if (c < NC) {
                float2 tc[32];
#pragma unroll
                for (int k = 0; k < 32; ++k) tc[k] = W[row(k) + c];                 // rows >= n are zero
#pragma unroll 1
                for (int r = 0; r < n; ++r) {
                    float ax = 0.f, ay = 0.f;
                    if (ga.real_roots) {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float rv = k < n ? __ldg(ga.Rr_root + (size_t)r * n + k).x : 0.f;                 // uniform
                            ax = fmaf(rv, tc[k].x, ax);
                            ay = fmaf(rv, tc[k].y, ay);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 32; ++k) {
                            const float2 rv = k < n ? __ldg(ga.Rr_root + (size_t)r * n + k) : make_float2(0.f, 0.f);    // uniform
                            ax = fmaf(rv.x, tc[k].x, fmaf(-rv.y, tc[k].y, ax));
                            ay = fmaf(rv.x, tc[k].y, fmaf(rv.y, tc[k].x, ay));
                        }
                    }
                    W[row(r) + c] = make_float2(ax, ay);
                }
            }
        }
        __syncwarp();
    }
    // ---- message (data.py:74-91): lane s < L draws section s
    int pos = 0, ks = 0;
    if (lane < g.L) {
        const uint4 w = philox4x32_10(make_uint4(flo, fhi, 2u, (unsigned)lane), key);
        pos = lane * g.M + (int)__umulhi(w.x, (unsigned)g.M);
        ks = (int)__umulhi(w.y, (unsigned)ga.K);
    }
    if (ga.x_out) {
        float2* xo = ga.x_out + f * g.N;
        for (int c = lane; c < g.N; c += 32) xo[c] = make_float2(0.f, 0.f);
        __syncwarp();
        if (lane < g.L) {
            xo[pos] = ga.sym[ks];
            ga.idx_out[f * g.L + lane] = (g.frame_base + f) * (long long)g.N + pos;
            ga.sym_out[f * g.L + lane] = ga.gray[ks];
        }
    }
    // ---- y = H x + w (channel.py:113-115), lane = row
    float2 yr = make_float2(0.f, 0.f);
    {
        const uint4 w = philox4x32_10(make_uint4(flo, fhi, 1u, (unsigned)(lane >> 1)), key);
        const float2 nz = (lane & 1) ? box_muller(w.z, w.w) : box_muller(w.x, w.y);
        yr = make_float2(nz.x * ga.noise_std, nz.y * ga.noise_std);
    }
    for (int s = 0; s < g.L; ++s) {
        const int ps = __shfl_sync(0xffffffffu, pos, s), kk = __shfl_sync(0xffffffffu, ks, s);
        const float2 hv = W[row(lane) + ps], sv = ga.sym[kk];
        yr.x = fmaf(hv.x, sv.x, fmaf(-hv.y, sv.y, yr.x));
        yr.y = fmaf(hv.x, sv.y, fmaf(hv.y, sv.x, yr.y));
    }
    yv[lane] = lane < n ? yr : make_float2(0.f, 0.f);
    __syncwarp();
}

}  // namespace ampsm
