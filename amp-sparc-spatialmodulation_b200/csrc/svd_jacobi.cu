// Batched thin SVD of wide complex64 matrices by one-sided (Hestenes) Jacobi: ONE WARP PER MATRIX, matrix in shared memory.
//
// Replaces the caller-side `torch.linalg.svd(A, full_matrices=False)` of the reference's VAMP driver
// (vamp_model.py:56-58) for per-frame channel matrices: H [n][N] (n <= 32 rows here, 33 .. 64 in the one-CTA-per-matrix kernel
// further down; n <= N) -> U [n][n], s [n] (descending),
// Vh [n][N] with H = U diag(s) Vh.  VAMP only ever uses V f(S) V^H and V S U^H (vamp.py:22,67,72), so the phases of the
// singular-vector pairs are free; they differ from LAPACK's.
//
// Method: the rows of W = H are orthogonalised in place by plane rotations (no Gram matrix: the condition number is
// not squared, float32 is enough also for correlated channels); the same rotations applied to Q = I give Q = U^H, so
// that W = U^H H = diag(s) Vh at convergence.  A sweep visits all n(n-1)/2 row pairs in 31 steps of 16 disjoint pairs.  Two
// lanes serve a pair -- lane (j,h) takes half of the columns of both rows (conflict-free 8-byte accesses, padded rows), forms
// the three inner products, combines them with its partner by one shuffle round and rotates in registers.  The order of the
// pairs is a recursive halving in which one row of every pair is STATIONARY for a whole phase and lives in registers (round 2b;
// the round-robin tournament of the first version moved both rows through shared memory every step and was bound by exactly
// that traffic: +34 % matrices/s).  A pair whose rows are already orthogonal to 4e-7 (relative) is skipped; the sweep loop ends
// when a whole sweep rotated nothing.  (Packed fp32x2 inner products and rotations -- half the FP instructions -- were measured
// after that change and bought nothing, +2 % / -5 % on the two entry points: the FMA pipe and the register file, not the issue
// slots, carry the 20 flops per element pair.)
#include <cstdlib>

#include "framegen.cuh"
#include "kernels.h"

namespace ampsm {

namespace {

constexpr int kSvdRows = 32;              // positions of the tournament (rows are zero-padded up to it)
constexpr int kSvdWarps = 3;              // warps (matrices) per CTA
constexpr int kSvdMaxSweeps = 16;
constexpr float kSvdTol = 4.0e-7f;

template <int NC, bool WITHY = false>
struct SvdShape {
    // Row placement of W.  NC = 64 (round 2b): rows 72 elements apart, row r shifted by rho(r) = (r & 7) ^ (r & 8 ? 7 : 0) elements --
    // every set of eight rows a half-warp touches in one access of the stationary-row ordering (a 3-dimensional subcube of the
    // row index: one of its low four bits fixed) then has eight distinct shifts, and with the two column halves 8 elements apart
    // the sixteen lanes hit sixteen distinct bank pairs.  (With rows NC + 1 apart the phases G <= 4 had two-way conflicts: 20 % of
    // all wavefronts of a kernel whose shared-memory pipe is 78 % busy.)  Other widths keep the odd stride.
    static constexpr bool swz = NC == 64 && WITHY;          // (with the n x n matrix Q next to W the wider rows would cost a CTA per SM)
    static constexpr int wstride = swz ? NC + 8 : NC + 1;     // complex elements per row of W
    static constexpr int qstride = kSvdRows + 1;
    static constexpr int q_elems = WITHY ? kSvdRows : kSvdRows * qstride;      // WITHY: the vector z = Q y instead of Q
    static constexpr int warp_bytes = (kSvdRows * wstride + (swz ? 8 : 0) + q_elems) * 8 + 2 * kSvdRows * 4;   // + perm, 1/s
};
template <bool SWZ>
__device__ __forceinline__ int svd_row(int r, int wstride) {
    return SWZ ? r * wstride + ((r & 7) ^ ((r & 8) ? 7 : 0)) : r * wstride;
}
// column k of a lane's share: lane half h takes the columns whose index mod 16 lies in [8h, 8h+8), so that the 16 lanes
// of a half-warp (8 consecutive rows x 2 halves, row stride = 1 mod 16 in 8-byte units) hit 16 distinct bank pairs
template <int NC>
__device__ __forceinline__ constexpr int svd_col(int c, int h) {
    return NC >= 16 ? ((c >> 3) * 16 + 8 * h + (c & 7)) : (2 * c + h);
}


// ---- block variant: four rows (two "super-rows") per group of four lanes -------------------------------------------------
// A step of the pairwise tournament loads and stores the whole matrix for 16 rotations; the kernel is bound by exactly that
// shared-memory traffic (ncu: data pipe 79 % busy).  The block variant runs the tournament over 16 super-rows of two rows:
// a group of four lanes loads the four rows of its super-row pair once (each lane a quarter of the columns, 128 registers
// at 64 columns), forms their 4 x 4 Gram matrix (two shuffle rounds), and works through the four cross pairs -- plus the two
// inner pairs once per sweep -- on the GRAM matrix alone, accumulating the plane rotations in a 4 x 4 matrix T that is then
// applied to the rows in one pass.  15 steps per sweep instead of 31: half the shared-memory traffic per sweep.
// (Correct -- the SVD tests pass with AMPSM_SVD_BLOCK=1 -- but slower on B200, see launch_svd_nc: an experiment kept as evidence.)
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ float2 conjf2(float2 a) { return make_float2(a.x, -a.y); }
// columns of lane h (0..3): inside every block of 16 columns the offsets {0,1,8,9}[h] + {0,2,4,6} -- with rows two apart in
// neighbouring groups (row stride = 1 mod 16 in 8-byte units) the 16 lanes of a half-warp hit 16 distinct bank pairs
template <int NC>
__device__ __forceinline__ constexpr int svd_bcol(int c, int h) {
    return NC >= 16 ? (16 * (c >> 2) + ((h & 1) + 8 * (h >> 1)) + 2 * (c & 3)) : (4 * c + h);
}
// one Jacobi rotation of rows (A, B) of the group, on the Gram matrix G (gd: diagonal, go[i][j] = <a_i, a_j> = sum conj(a_i) a_j)
// and on the accumulated transformation T (rows' = T rows); same formulas as the pairwise kernel
template <int A, int B>
__device__ __forceinline__ bool svd_grot(float (&gd)[4], float2 (&go)[4][4], float2 (&T)[4][4]) {
    const float alpha = gd[A], beta = gd[B];
    const float2 gm = go[A][B];
    const float g2 = gm.x * gm.x + gm.y * gm.y;
    if (!(g2 > kSvdTol * kSvdTol * alpha * beta && g2 > 0.f)) return false;
    const float gabs = sqrtf(g2), ginv = 1.0f / gabs;
    const float pr = gm.x * ginv, pi = -gm.y * ginv;                                 // e^{-i phi} = conj(gamma) / |gamma|
    const float zeta = (beta - alpha) * (0.5f * ginv);
    const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
    const float cs = rsqrtf(fmaf(t, t, 1.0f)), sn = cs * t;
    const float2 sp = make_float2(sn * pr, sn * pi), cp = make_float2(cs * pr, cs * pi);
    // a' = c a - (s p) b ; b' = s a + (c p) b
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float2 ta = T[A][k], tb = T[B][k], u = cmulf(sp, tb), v = cmulf(cp, tb);
        T[A][k] = make_float2(cs * ta.x - u.x, cs * ta.y - u.y);
        T[B][k] = make_float2(sn * ta.x + v.x, sn * ta.y + v.y);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k == A || k == B) continue;
        const float2 x = go[k][A], y = go[k][B], u = cmulf(sp, y), v = cmulf(cp, y);   // <a_k, .> is linear in its second argument
        go[k][A] = make_float2(cs * x.x - u.x, cs * x.y - u.y);
        go[k][B] = make_float2(sn * x.x + v.x, sn * x.y + v.y);
        go[A][k] = conjf2(go[k][A]);
        go[B][k] = conjf2(go[k][B]);
    }
    gd[A] = alpha - t * gabs;
    gd[B] = beta + t * gabs;
    go[A][B] = go[B][A] = make_float2(0.f, 0.f);
    return true;
}

// NC = number of columns (compile time), n = number of rows (run time, <= 32).
// WITHY: the caller only needs U^H y (VAMP's y~ = diag(s) U^H y, vamp.py:22): the rotations are applied to the vector y
// instead of the n x n matrix Q = U^H -- a third less shared-memory traffic per rotation, no U written or read later.
// BLOCK: the four-rows-per-group sweep above instead of the pairwise one.
// GEN: the frame is drawn inside the kernel (framegen.cuh) straight into W and z -- H never exists in HBM (WITHY only).
struct SvdGen {
    GenArgs ga;
    Geom g;
};
template <int NC, bool WITHY, bool BLOCK, bool GEN = false>
__global__ void __launch_bounds__(kSvdWarps * 32, BLOCK ? 1 : (WITHY ? 4 : 3)) svd_jacobi_kernel(const float2* __restrict__ H, long long frames, int n, float2* __restrict__ U,
                                                                   float* __restrict__ S, float2* __restrict__ Vh, int* __restrict__ sweeps_out,
                                                                   const float2* __restrict__ yin, float2* __restrict__ yrot,
                                                                   const __grid_constant__ SvdGen gen) {
    static_assert(!GEN || WITHY, "in-kernel generation feeds the fused VAMP path");
    using Sh = SvdShape<NC, WITHY>;
    constexpr int HC = NC / 2, HQ = kSvdRows / 2;             // columns of W / Q per lane
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    float2* W = reinterpret_cast<float2*>(smem + (size_t)wic * Sh::warp_bytes);
    float2* Q = W + kSvdRows * Sh::wstride + (Sh::swz ? 8 : 0);      // the shifted last row ends up to 7 elements later
    int* perm = reinterpret_cast<int*>(Q + Sh::q_elems);
    float* sinv = reinterpret_cast<float*>(perm + kSvdRows);
    float* nrm = sinv;                                        // squared row norms during the sweeps
    const int j = lane >> 1, h = lane & 1;

    for (long long f = (long long)blockIdx.x * kSvdWarps + wic; f < frames; f += (long long)gridDim.x * kSvdWarps) {
        // ---- load: W = H (rows >= n are zero), Q = I
        if constexpr (GEN) {
            gen_frame<NC, Sh::swz>(gen.ga, gen.g, f, n, W, Sh::wstride, Q, lane);
        } else {
            const float2* Hf = H + f * (long long)n * NC;
            for (int e = lane; e < kSvdRows * NC; e += 32) {
                const int r = e / NC, c = e - r * NC;
                W[svd_row<Sh::swz>(r, Sh::wstride) + c] = r < n ? __ldg(Hf + e) : make_float2(0.f, 0.f);
            }
        }
        if constexpr (GEN) {
        } else if constexpr (WITHY) {
            Q[lane] = lane < n ? __ldg(yin + f * n + lane) : make_float2(0.f, 0.f);      // z = y (rows >= n are zero)
        } else {
            for (int e = lane; e < kSvdRows * kSvdRows; e += 32) {
                const int r = e >> 5, c = e & 31;
                Q[r * Sh::qstride + c] = make_float2(r == c ? 1.f : 0.f, 0.f);
            }
        }
        __syncwarp();

        int sweeps = 0;
        int clean = 0;                                         // consecutive steps in which no pair of the warp rotated
        bool done = false;
        for (; sweeps < kSvdMaxSweeps; ++sweeps) {
            bool rotated = false;
            if constexpr (BLOCK) {
                constexpr int QC = NC / 4, kSuper = kSvdRows / 2;        // columns per lane, super-rows
                const int g = lane >> 2, hb = lane & 3;
                for (int step = 0; step < kSuper - 1; ++step) {
                    int P, Qs;
                    if (g == 0) {
                        P = kSuper - 1;
                        Qs = step;
                    } else {
                        P = step + g;
                        if (P >= kSuper - 1) P -= kSuper - 1;
                        Qs = step - g;
                        if (Qs < 0) Qs += kSuper - 1;
                    }
                    const int rows[4] = {2 * P, 2 * P + 1, 2 * Qs, 2 * Qs + 1};
                    float2 a[4][QC];
                    float gd[4] = {0.f, 0.f, 0.f, 0.f};
                    float2 go[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 4; ++k) go[i][k] = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int c = 0; c < QC; ++c) a[i][c] = W[svd_row<Sh::swz>(rows[i], Sh::wstride) + svd_bcol<NC>(c, hb)];
#pragma unroll
                    for (int c = 0; c < QC; ++c) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            gd[i] = fmaf(a[i][c].x, a[i][c].x, fmaf(a[i][c].y, a[i][c].y, gd[i]));
#pragma unroll
                            for (int k = i + 1; k < 4; ++k) {
                                go[i][k].x = fmaf(a[i][c].x, a[k][c].x, fmaf(a[i][c].y, a[k][c].y, go[i][k].x));      // sum conj(a_i) a_k
                                go[i][k].y = fmaf(a[i][c].x, a[k][c].y, fmaf(-a[i][c].y, a[k][c].x, go[i][k].y));
                            }
                        }
                    }
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            gd[i] += __shfl_xor_sync(0xffffffffu, gd[i], o);
#pragma unroll
                            for (int k = i + 1; k < 4; ++k) {
                                go[i][k].x += __shfl_xor_sync(0xffffffffu, go[i][k].x, o);
                                go[i][k].y += __shfl_xor_sync(0xffffffffu, go[i][k].y, o);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = i + 1; k < 4; ++k) go[k][i] = conjf2(go[i][k]);
                    float2 T[4][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int k = 0; k < 4; ++k) T[i][k] = make_float2(i == k ? 1.f : 0.f, 0.f);
                    bool any = false;
                    any |= svd_grot<0, 2>(gd, go, T);
                    any |= svd_grot<1, 3>(gd, go, T);
                    any |= svd_grot<0, 3>(gd, go, T);
                    any |= svd_grot<1, 2>(gd, go, T);
                    if (step == 0) {            // the pairs inside a super-row: once per sweep
                        any |= svd_grot<0, 1>(gd, go, T);
                        any |= svd_grot<2, 3>(gd, go, T);
                    }
                    if (any) {                  // uniform over the four lanes of the group (they hold the same Gram matrix)
#pragma unroll
                        for (int c = 0; c < QC; ++c) {
                            const float2 x0 = a[0][c], x1 = a[1][c], x2 = a[2][c], x3 = a[3][c];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                float2 r = cmulf(T[i][0], x0);
                                const float2 r1 = cmulf(T[i][1], x1), r2 = cmulf(T[i][2], x2), r3 = cmulf(T[i][3], x3);
                                r = make_float2((r.x + r1.x) + (r2.x + r3.x), (r.y + r1.y) + (r2.y + r3.y));
                                W[svd_row<Sh::swz>(rows[i], Sh::wstride) + svd_bcol<NC>(c, hb)] = r;
                            }
                        }
                        if constexpr (WITHY) {
                            if (hb == 0) {
                                const float2 z0 = Q[rows[0]], z1 = Q[rows[1]], z2 = Q[rows[2]], z3 = Q[rows[3]];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 r0 = cmulf(T[i][0], z0), r1 = cmulf(T[i][1], z1), r2 = cmulf(T[i][2], z2), r3 = cmulf(T[i][3], z3);
                                    Q[rows[i]] = make_float2((r0.x + r1.x) + (r2.x + r3.x), (r0.y + r1.y) + (r2.y + r3.y));
                                }
                            }
                        } else {
#pragma unroll
                            for (int c = 0; c < kSvdRows / 4; ++c) {
                                const int cq = svd_bcol<kSvdRows>(c, hb);
                                const float2 x0 = Q[rows[0] * Sh::qstride + cq], x1 = Q[rows[1] * Sh::qstride + cq],
                                             x2 = Q[rows[2] * Sh::qstride + cq], x3 = Q[rows[3] * Sh::qstride + cq];
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 r0 = cmulf(T[i][0], x0), r1 = cmulf(T[i][1], x1), r2 = cmulf(T[i][2], x2), r3 = cmulf(T[i][3], x3);
                                    Q[rows[i] * Sh::qstride + cq] = make_float2((r0.x + r1.x) + (r2.x + r3.x), (r0.y + r1.y) + (r2.y + r3.y));
                                }
                            }
                        }
                        rotated = true;
                    }
                    __syncwarp();
                }
            } else {
            // Pair ordering: recursive halving with a STATIONARY row.  Phase G = 16, 8, 4, 2, 1: the 32 rows form super-blocks of
            // 2G rows; lane pair j = (sb, i) keeps row p = 2G sb + i of the first half in REGISTERS for the whole phase and meets
            // the rows q = 2G sb + G + (i + s) mod G of the second half, s = 0 .. G-1 -- 16 + 8 + 4 + 2 + 1 = 31 steps of 16
            // disjoint pairs that cover every pair once, like the round-robin tournament of the first version, but only ONE row
            // of a pair crosses shared memory per step (the kernel is bound by exactly that traffic: ncu, data pipe 91 % busy).
            // Neighbouring lane pairs still touch rows that differ mod 16, so the accesses stay conflict-free.
            // The squared row norms are kept in `nrm` (the 1/s scratch, free until the output stage): exact at the start of every
            // sweep, then carried through the rotations (alpha' = alpha - t |gamma|, beta' = beta + t |gamma|).  Only gamma is
            // formed from the rows: 4 instead of 8 FMAs per element pair, half the work of a pair that does not rotate.  The
            // carried norms drift by a few ulps per rotation (<= 31 per sweep); they enter the skip test and the rotation angle
            // only -- convergence is decided by the exact gamma, the singular values by exact norms at the end.
            {
                float nn = 0.f;
#pragma unroll 8
                for (int c = 0; c < NC; ++c) {
                    const float2 wv = W[svd_row<Sh::swz>(lane, Sh::wstride) + c];
                    nn = fmaf(wv.x, wv.x, fmaf(wv.y, wv.y, nn));
                }
                nrm[lane] = nn;
                __syncwarp();
            }
            for (int G = kSvdRows / 2; G >= 1; G >>= 1) {
                const int sb = j / G, i = j - sb * G;
                const int p = sb * 2 * G + i;
                float2* wa = W + svd_row<Sh::swz>(p, Sh::wstride);
                float2 a[HC];
#pragma unroll
                for (int c = 0; c < HC; ++c) a[c] = wa[svd_col<NC>(c, h)];
                float alpha = nrm[p];
                bool a_dirty = false;
                for (int s = 0; s < G; ++s) {
                    const int q = sb * 2 * G + G + (i ^ s);              // XOR, not a rotation: every access set is a subcube of the row index
                    float2* wb = W + svd_row<Sh::swz>(q, Sh::wstride);
                    float2 b[HC];
                    float gr = 0.f, gi = 0.f;
                    const float beta = nrm[q];
#pragma unroll
                    for (int c = 0; c < HC; ++c) {
                        b[c] = wb[svd_col<NC>(c, h)];
                        gr = fmaf(a[c].x, b[c].x, fmaf(a[c].y, b[c].y, gr));          // gamma = sum conj(a) b
                        gi = fmaf(a[c].x, b[c].y, fmaf(-a[c].y, b[c].x, gi));
                    }
                    gr += __shfl_xor_sync(0xffffffffu, gr, 1);
                    gi += __shfl_xor_sync(0xffffffffu, gi, 1);
                    const float g2 = gr * gr + gi * gi;
                    const bool rot = g2 > kSvdTol * kSvdTol * alpha * beta && g2 > 0.f;
                    clean = __any_sync(0xffffffffu, rot) ? 0 : clean + 1;
                    if (rot) {
                        // b~ = e^{-i phi} b makes <a, b~> = |gamma| real; then the real Jacobi rotation:
                        // zeta = (beta - alpha) / (2 |gamma|), t = sign(zeta) / (|zeta| + sqrt(1 + zeta^2)), c = 1/sqrt(1+t^2), s = c t
                        const float gabs = sqrtf(g2), ginv = 1.0f / gabs;
                        const float pr = gr * ginv, pi = -gi * ginv;                    // e^{-i phi} = conj(gamma) / |gamma|
                        const float zeta = (beta - alpha) * (0.5f * ginv);
                        const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
                        const float cs = rsqrtf(fmaf(t, t, 1.0f)), sn = cs * t;
                        const float spr = sn * pr, spi = sn * pi, cpr = cs * pr, cpi = cs * pi;
                        alpha = fmaf(-t, gabs, alpha);
                        if (h == 0) nrm[q] = fmaf(t, gabs, beta);
                        // a' = c a - (s p) b ; b' = s a + (c p) b
#pragma unroll
                        for (int c = 0; c < HC; ++c) {
                            const float2 x = a[c], y = b[c];
                            a[c] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                            wb[svd_col<NC>(c, h)] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                        }
                        a_dirty = true;
                        if constexpr (WITHY) {
                            if (h == 0) {           // the same rotation on the two entries of z = Q y
                                const float2 x = Q[p], y = Q[q];
                                Q[p] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                                Q[q] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                            }
                        } else {
                            // (keeping this row of Q in registers as well was measured: 252 registers or spills, slower)
                            float2* qa = Q + p * Sh::qstride;
                            float2* qb = Q + q * Sh::qstride;
#pragma unroll
                            for (int c = 0; c < HQ; ++c) {
                                const int cq = svd_col<kSvdRows>(c, h);
                                const float2 x = qa[cq], y = qb[cq];
                                qa[cq] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                                qb[cq] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                            }
                        }
                        rotated = true;
                    }
                    __syncwarp();
                }
                if (a_dirty) {
#pragma unroll
                    for (int c = 0; c < HC; ++c) wa[svd_col<NC>(c, h)] = a[c];
                    if (h == 0) nrm[p] = alpha;
                }
                __syncwarp();
                // any 31 consecutive steps of the schedule visit every pair once: 31 steps without a rotation anywhere in the warp
                // certify the matrix, wherever in a sweep they end (checked between phases, when the stationary rows are home)
                if (clean >= kSvdRows - 1) {
                    done = true;
                    break;
                }
            }
            }
            if (done || !__any_sync(0xffffffffu, rotated)) {
                ++sweeps;
                break;
            }
        }

        // ---- singular values = row norms of W (lane r owns row r), descending order by rank counting
        float rn2 = 0.f;
#pragma unroll 8
        for (int c = 0; c < NC; ++c) {
            const float2 w = W[svd_row<Sh::swz>(lane, Sh::wstride) + c];
            rn2 = fmaf(w.x, w.x, fmaf(w.y, w.y, rn2));
        }
        const float sv = sqrtf(rn2);
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
                const float2 w = W[svd_row<Sh::swz>(row, Sh::wstride) + c];
                out[c] = make_float2(w.x * sc, w.y * sc);
            }
        }
        if constexpr (WITHY) {
            if (lane < n) yrot[f * n + lane] = Q[perm[lane]];                   // (U^H y)_k, k in singular-value order
        } else if (lane < n) {
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

template <int NC, bool WITHY, bool BLOCK>
int launch_svd_nc_b(const float2* H, long long frames, int n, float2* U, float* S, float2* Vh, int* sweeps, const float2* y, float2* yrot,
                  cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = svd_jacobi_kernel<NC, WITHY, BLOCK>;
    const size_t smem = (size_t)SvdShape<NC, WITHY>::warp_bytes * kSvdWarps;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(svd)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSvdWarps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (frames + kSvdWarps - 1) / kSvdWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kSvdWarps * 32, smem, stream>>>(H, frames, n, U, S, Vh, sweeps, y, yrot, SvdGen{});
    count_launch();
    return check_cuda(cudaGetLastError(), "svd_jacobi_kernel launch");
}

template <int NC>
int launch_svd_gen_nc(const GenArgs& ga, const Geom& g, long long frames, float* S, float2* Vh, float2* yrot, cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = svd_jacobi_kernel<NC, true, false, true>;
    const size_t smem = (size_t)SvdShape<NC, true>::warp_bytes * kSvdWarps;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(svd gen)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSvdWarps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (frames + kSvdWarps - 1) / kSvdWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    SvdGen sg{ga, g};
    kern<<<(unsigned)grid, kSvdWarps * 32, smem, stream>>>(nullptr, frames, g.n, nullptr, S, Vh, nullptr, nullptr, yrot, sg);
    count_launch();
    return check_cuda(cudaGetLastError(), "svd_jacobi_kernel (generated frames) launch");
}

// the frames of the same stream written out (the checker of the fused path, and one-pass input generation for the detectors
// that read H from HBM): one warp per frame, tile in shared memory, coalesced stores
template <int NC>
__global__ void __launch_bounds__(kSvdWarps * 32) generate_frames_kernel(const __grid_constant__ SvdGen gen, long long frames, float2* __restrict__ H,
                                                                        float2* __restrict__ y) {
    constexpr int wstride = NC + 1;
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, wic = threadIdx.x >> 5;
    float2* W = reinterpret_cast<float2*>(smem) + (size_t)wic * (kSvdRows * wstride + kSvdRows);
    float2* yv = W + kSvdRows * wstride;
    const int n = gen.g.n;
    for (long long f = (long long)blockIdx.x * kSvdWarps + wic; f < frames; f += (long long)gridDim.x * kSvdWarps) {
        gen_frame<NC, false>(gen.ga, gen.g, f, n, W, wstride, yv, lane);
        if (H)
            for (int e = lane; e < n * NC; e += 32) {
                const int r = e / NC, c = e - r * NC;
                H[f * (long long)n * NC + e] = W[r * wstride + c];
            }
        if (y && lane < n) y[f * n + lane] = yv[lane];
        __syncwarp();
    }
}

template <int NC>
int launch_generate_nc(const GenArgs& ga, const Geom& g, long long frames, float2* H, float2* y, cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = generate_frames_kernel<NC>;
    const size_t smem = (size_t)(kSvdRows * (NC + 1) + kSvdRows) * 8 * kSvdWarps;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(generate)"))
        return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSvdWarps * 32, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    const long long need = (frames + kSvdWarps - 1) / kSvdWarps;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    SvdGen sg{ga, g};
    kern<<<(unsigned)grid, kSvdWarps * 32, smem, stream>>>(sg, frames, H, y);
    count_launch();
    return check_cuda(cudaGetLastError(), "generate_frames_kernel launch");
}

template <int NC, bool WITHY>
int launch_svd_nc(const float2* H, long long frames, int n, float2* U, float* S, float2* Vh, int* sweeps, const float2* y, float2* yrot,
                  cudaStream_t stream) {
    // measured on B200 (262 144 matrices 32 x 64, from-H path): pairwise 3.59e6 frames/s, block 1.82e6 -- the block sweep halves
    // the shared-memory traffic but needs 255 registers (6 warps per SM instead of 12) and ~6x the serial scalar work per
    // step; it stays selectable for A/B runs
    static const bool block = getenv("AMPSM_SVD_BLOCK") != nullptr;
    return block ? launch_svd_nc_b<NC, WITHY, true>(H, frames, n, U, S, Vh, sweeps, y, yrot, stream)
                 : launch_svd_nc_b<NC, WITHY, false>(H, frames, n, U, S, Vh, sweeps, y, yrot, stream);
}


// ---- 33 .. 64 rows (BASELINE config 3: 64 x 128): ONE CTA OF 128 THREADS PER MATRIX --------------------------------------------
// Same method, 64 tournament positions: 63 steps of 32 disjoint pairs per sweep, FOUR lanes per pair -- lane (j, h) takes the
// columns 16 (c >> 2) + 4 h + (c & 3) of both rows (with rows one apart in neighbouring pairs and an odd row stride the 16 lanes
// of a half-warp hit 16 distinct bank pairs), forms its share of the three inner products, combines them with its three partners
// by two shuffle rounds and rotates in registers.  The pairs of a step are disjoint, so the four warps only meet at one CTA
// barrier per step.  W (64 x 129 complex64 = 66 KiB) + Q (64 x 65) or the rotated vector y live in shared memory.
constexpr int kSvd64Rows = 64, kSvd64Threads = 128, kSvd64MaxSweeps = 20;
template <int NC, bool WITHY>
struct Svd64Shape {
    static constexpr int wstride = NC + 1;
    static constexpr int qstride = kSvd64Rows + 1;
    static constexpr int q_elems = WITHY ? kSvd64Rows : kSvd64Rows * qstride;
    static constexpr int bytes = (kSvd64Rows * wstride + q_elems) * 8 + 2 * kSvd64Rows * 4 + 16;   // + perm, 1/s, flags
};
template <int NC>
__device__ __forceinline__ constexpr int svd64_col(int c, int h) { return 16 * (c >> 2) + 4 * h + (c & 3); }

template <int NC, bool WITHY>
__global__ void __launch_bounds__(kSvd64Threads) svd_jacobi64_kernel(const float2* __restrict__ H, long long frames, int n, float2* __restrict__ U,
                                                                    float* __restrict__ S, float2* __restrict__ Vh, int* __restrict__ sweeps_out,
                                                                    const float2* __restrict__ yin, float2* __restrict__ yrot) {
    using Sh = Svd64Shape<NC, WITHY>;
    constexpr int QC = NC / 4, QQ = kSvd64Rows / 4;            // columns of W / Q per lane
    extern __shared__ __align__(16) unsigned char smem[];
    float2* W = reinterpret_cast<float2*>(smem);
    float2* Q = W + kSvd64Rows * Sh::wstride;
    int* perm = reinterpret_cast<int*>(Q + Sh::q_elems);
    float* sinv = reinterpret_cast<float*>(perm + kSvd64Rows);
    float* nrm = sinv;                                         // squared row norms during the sweeps
    const int tid = threadIdx.x;
    const int j = tid >> 2, h = tid & 3;

    for (long long f = blockIdx.x; f < frames; f += gridDim.x) {
        const float2* Hf = H + f * (long long)n * NC;
        for (int e = tid; e < kSvd64Rows * NC; e += kSvd64Threads) {
            const int r = e / NC, c = e - r * NC;
            W[r * Sh::wstride + c] = r < n ? __ldg(Hf + e) : make_float2(0.f, 0.f);
        }
        if constexpr (WITHY) {
            if (tid < kSvd64Rows) Q[tid] = tid < n ? __ldg(yin + f * n + tid) : make_float2(0.f, 0.f);
        } else {
            for (int e = tid; e < kSvd64Rows * kSvd64Rows; e += kSvd64Threads) {
                const int r = e >> 6, c = e & 63;
                Q[r * Sh::qstride + c] = make_float2(r == c ? 1.f : 0.f, 0.f);
            }
        }
        __syncthreads();

        int sweeps = 0;
        for (; sweeps < kSvd64MaxSweeps; ++sweeps) {
            bool rotated = false;
            // the stationary-row ordering of the one-warp kernel: phases G = 32, 16, ..., 1; thread group j = (sb, i) keeps row
            // p = 2G sb + i in registers for the phase and meets q = 2G sb + G + (i + s) mod G, s = 0 .. G-1 (63 steps per sweep)
            {   // squared row norms, exact at the start of the sweep and carried through its rotations (see the one-warp kernel)
                const int r = tid >> 1, half = tid & 1;
                float nn = 0.f;
                for (int c = half; c < NC; c += 2) {
                    const float2 wv = W[r * Sh::wstride + c];
                    nn = fmaf(wv.x, wv.x, fmaf(wv.y, wv.y, nn));
                }
                nn += __shfl_xor_sync(0xffffffffu, nn, 1);
                if (half == 0) nrm[r] = nn;
            }
            __syncthreads();
            for (int G = kSvd64Rows / 2; G >= 1; G >>= 1) {
                const int sb = j / G, i = j - sb * G;
                const int p = sb * 2 * G + i;
                float2* wa = W + p * Sh::wstride;
                float2 a[QC];
#pragma unroll
                for (int c = 0; c < QC; ++c) a[c] = wa[svd64_col<NC>(c, h)];
                float alpha = nrm[p];
                bool a_dirty = false;
                for (int s = 0; s < G; ++s) {
                    const int q = sb * 2 * G + G + (i ^ s);
                    float2* wb = W + q * Sh::wstride;
                    float2 b[QC];
                    float gr = 0.f, gi = 0.f;
                    const float beta = nrm[q];
#pragma unroll
                    for (int c = 0; c < QC; ++c) {
                        b[c] = wb[svd64_col<NC>(c, h)];
                        gr = fmaf(a[c].x, b[c].x, fmaf(a[c].y, b[c].y, gr));          // gamma = sum conj(a) b
                        gi = fmaf(a[c].x, b[c].y, fmaf(-a[c].y, b[c].x, gi));
                    }
#pragma unroll
                    for (int o = 1; o <= 2; o <<= 1) {
                        gr += __shfl_xor_sync(0xffffffffu, gr, o);
                        gi += __shfl_xor_sync(0xffffffffu, gi, o);
                    }
                    const float g2 = gr * gr + gi * gi;
                    const bool rot = g2 > kSvdTol * kSvdTol * alpha * beta && g2 > 0.f;
                    if (rot) {
                        const float gabs = sqrtf(g2), ginv = 1.0f / gabs;
                        const float pr = gr * ginv, pi = -gi * ginv;                    // e^{-i phi} = conj(gamma) / |gamma|
                        const float zeta = (beta - alpha) * (0.5f * ginv);
                        const float t = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(fmaf(zeta, zeta, 1.0f)));
                        const float cs = rsqrtf(fmaf(t, t, 1.0f)), sn = cs * t;
                        const float spr = sn * pr, spi = sn * pi, cpr = cs * pr, cpi = cs * pi;
                        alpha = fmaf(-t, gabs, alpha);
                        if (h == 0) nrm[q] = fmaf(t, gabs, beta);
                        // a' = c a - (s p) b ; b' = s a + (c p) b
#pragma unroll
                        for (int c = 0; c < QC; ++c) {
                            const float2 x = a[c], y = b[c];
                            a[c] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                            wb[svd64_col<NC>(c, h)] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                        }
                        a_dirty = true;
                        if constexpr (WITHY) {
                            if (h == 0) {
                                const float2 x = Q[p], y = Q[q];
                                Q[p] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                                Q[q] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                            }
                        } else {
                            float2* qa = Q + p * Sh::qstride;
                            float2* qb = Q + q * Sh::qstride;
#pragma unroll
                            for (int c = 0; c < QQ; ++c) {
                                const int cq = svd64_col<kSvd64Rows>(c, h);
                                const float2 x = qa[cq], y = qb[cq];
                                qa[cq] = make_float2(fmaf(cs, x.x, fmaf(-spr, y.x, spi * y.y)), fmaf(cs, x.y, fmaf(-spr, y.y, -spi * y.x)));
                                qb[cq] = make_float2(fmaf(sn, x.x, fmaf(cpr, y.x, -cpi * y.y)), fmaf(sn, x.y, fmaf(cpr, y.y, cpi * y.x)));
                            }
                        }
                        rotated = true;
                    }
                    __syncthreads();
                }
                if (a_dirty) {
#pragma unroll
                    for (int c = 0; c < QC; ++c) wa[svd64_col<NC>(c, h)] = a[c];
                    if (h == 0) nrm[p] = alpha;
                }
                __syncthreads();
            }
            if (!__syncthreads_or(rotated ? 1 : 0)) {
                ++sweeps;
                break;
            }
        }

        // ---- singular values = row norms of W (two threads per row), descending order by rank counting
        {
            const int r = tid >> 1, half = tid & 1;
            float nrm = 0.f;
            for (int c = half; c < NC; c += 2) {
                const float2 w = W[r * Sh::wstride + c];
                nrm = fmaf(w.x, w.x, fmaf(w.y, w.y, nrm));
            }
            nrm += __shfl_xor_sync(0xffffffffu, nrm, 1);
            if (half == 0) sinv[r] = sqrtf(nrm);                       // sinv holds s for the moment
        }
        __syncthreads();
        float sv = 0.f;
        int rank = 0;
        if (tid < kSvd64Rows) {
            sv = sinv[tid];
            for (int o = 0; o < kSvd64Rows; ++o) {
                const float other = sinv[o];
                rank += (other > sv) || (other == sv && o < tid);
            }
        }
        __syncthreads();
        if (tid < kSvd64Rows) {
            perm[rank] = tid;
            sinv[rank] = sv > 0.f ? 1.0f / sv : 0.f;
            if (rank < n) S[f * n + rank] = sv;
        }
        __syncthreads();
        // Vh[k][:] = W[row_k][:] / s_k, U[i][k] = conj(Q[row_k][i]): coalesced stores
        for (int e = tid; e < n * NC; e += kSvd64Threads) {
            const int k = e / NC, c = e - k * NC;
            const float2 w = W[perm[k] * Sh::wstride + c];
            const float sc = sinv[k];
            Vh[(f * n + k) * (long long)NC + c] = make_float2(w.x * sc, w.y * sc);
        }
        if constexpr (WITHY) {
            if (tid < n) yrot[f * n + tid] = Q[perm[tid]];
        } else {
            for (int e = tid; e < n * n; e += kSvd64Threads) {
                const int i = e / n, k = e - i * n;
                const float2 qv = Q[perm[k] * Sh::qstride + i];
                U[(f * n + i) * (long long)n + k] = make_float2(qv.x, -qv.y);
            }
        }
        if (sweeps_out && tid == 0) sweeps_out[f] = sweeps;
        __syncthreads();
    }
}

template <int NC, bool WITHY>
int launch_svd64(const float2* H, long long frames, int n, float2* U, float* S, float2* Vh, int* sweeps, const float2* y, float2* yrot,
                 cudaStream_t stream) {
    int dev = 0, sms = 0;
    if (int e = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return e;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    auto kern = svd_jacobi64_kernel<NC, WITHY>;
    const size_t smem = Svd64Shape<NC, WITHY>::bytes;
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(svd64)")) return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kSvd64Threads, smem);
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sms * per_sm;
    if (grid > frames) grid = frames;
    if (grid < 1) grid = 1;
    kern<<<(unsigned)grid, kSvd64Threads, smem, stream>>>(H, frames, n, U, S, Vh, sweeps, y, yrot);
    count_launch();
    return check_cuda(cudaGetLastError(), "svd_jacobi64_kernel launch");
}

__global__ void identity_kernel(float2* I, int n) {
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) I[e] = make_float2((e / n) == (e % n) ? 1.f : 0.f, 0.f);
}

}  // namespace

// y == nullptr: U, s, Vh.  y != nullptr: s, Vh and yrot = U^H y (U is not formed; see WITHY).
int launch_svd_jacobi(const float2* H, long long frames, int n, int N, float2* U, float* S, float2* Vh, int* sweeps, const float2* y,
                      float2* yrot, cudaStream_t stream) {
    if (n < 1 || n > kSvd64Rows || N < n) {
        set_error("batched SVD: needs 1 <= n <= 64 rows and n <= N columns (got %d x %d)", n, N);
        return AMPSM_ENOFIT;
    }
    if (n > kSvdRows) {                                        // 33 .. 64 rows: one CTA per matrix
        if (N == 64) return y ? launch_svd64<64, true>(H, frames, n, U, S, Vh, sweeps, y, yrot, stream)
                              : launch_svd64<64, false>(H, frames, n, U, S, Vh, sweeps, nullptr, nullptr, stream);
        if (N == 128) return y ? launch_svd64<128, true>(H, frames, n, U, S, Vh, sweeps, y, yrot, stream)
                               : launch_svd64<128, false>(H, frames, n, U, S, Vh, sweeps, nullptr, nullptr, stream);
        set_error("batched SVD with more than 32 rows: column counts 64 and 128 are instantiated (got %d)", N);
        return AMPSM_ENOFIT;
    }
#define AMPSM_SVD_NC(NCC) \
    case NCC: return y ? launch_svd_nc<NCC, true>(H, frames, n, U, S, Vh, sweeps, y, yrot, stream) \
                       : launch_svd_nc<NCC, false>(H, frames, n, U, S, Vh, sweeps, nullptr, nullptr, stream);
    switch (N) {
        AMPSM_SVD_NC(8)
        AMPSM_SVD_NC(16)
        AMPSM_SVD_NC(32)
        AMPSM_SVD_NC(64)
        default:
            set_error("batched SVD: column counts 8, 16, 32, 64 are instantiated (got %d)", N);
            return AMPSM_ENOFIT;
    }
#undef AMPSM_SVD_NC
}

static int check_gen_shape(const Geom& g) {
    if (g.n < 1 || g.n > kSvdRows || g.N < g.n || g.L < 1 || g.L > 32 || g.Lin != 1 || g.M * g.L != g.N) {
        set_error("frame generation: needs 1 <= n <= 32 rows, n <= N, Lin = 1 and at most 32 sections (got n=%d N=%d L=%d Lin=%d)", g.n, g.N,
                  g.L, g.Lin);
        return AMPSM_ENOFIT;
    }
    return 0;
}

int launch_svd_jacobi_gen(const GenArgs& gen, const Geom& g, long long frames, float* S, float2* Vh, float2* yrot, cudaStream_t stream) {
    if (int e = check_gen_shape(g)) return e;
    switch (g.N) {
        case 8: return launch_svd_gen_nc<8>(gen, g, frames, S, Vh, yrot, stream);
        case 16: return launch_svd_gen_nc<16>(gen, g, frames, S, Vh, yrot, stream);
        case 32: return launch_svd_gen_nc<32>(gen, g, frames, S, Vh, yrot, stream);
        case 64: return launch_svd_gen_nc<64>(gen, g, frames, S, Vh, yrot, stream);
        default:
            set_error("frame generation: column counts 8, 16, 32, 64 are instantiated (got %d)", g.N);
            return AMPSM_ENOFIT;
    }
}

int launch_generate_frames(const GenArgs& gen, const Geom& g, long long frames, float2* H, float2* y, cudaStream_t stream) {
    if (int e = check_gen_shape(g)) return e;
    switch (g.N) {
        case 8: return launch_generate_nc<8>(gen, g, frames, H, y, stream);
        case 16: return launch_generate_nc<16>(gen, g, frames, H, y, stream);
        case 32: return launch_generate_nc<32>(gen, g, frames, H, y, stream);
        case 64: return launch_generate_nc<64>(gen, g, frames, H, y, stream);
        default:
            set_error("frame generation: column counts 8, 16, 32, 64 are instantiated (got %d)", g.N);
            return AMPSM_ENOFIT;
    }
}

int launch_identity(float2* I, int n, cudaStream_t stream) {
    identity_kernel<<<1, 256, 0, stream>>>(I, n);
    count_launch();
    return check_cuda(cudaGetLastError(), "identity_kernel launch");
}

}  // namespace ampsm
