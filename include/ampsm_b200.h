/*
 * ampsm_b200.h -- C ABI of libampsm_b200.so: B200 (sm_100a) kernels for the BAMP / SCAMP / VAMP
 * spatial-modulation detector hot path of AhmedKishki/AMP-SPARC-SpatialModulation.
 *
 * The reference has no FFI: the path sits behind Python nn.Module objects.  Each entry
 * point below names the reference interface it replaces; the Python classes in
 * amp-sparc-spatialmodulation_b200/{bamp,vamp,scamp,loss}.py bind these symbols with ctypes and keep the
 * reference's call signatures.  INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - complex64 = interleaved float {re,im}; complex128 = interleaved double; matrices row-major.
 *   - "frame" = one reference call with batch=1 (per-frame soft-max shift, exit test and pooled variance).
 *   - *_detect  : every pointer is DEVICE memory, work is enqueued on `stream` (a cudaStream_t) and the call
 *                 returns without synchronising; the caller owns all buffers; calls on different streams may run
 *                 concurrently.  Optional outputs may be NULL.
 *   - *_detect_host : every pointer is HOST memory; the call copies inputs to the device in chunks (copies
 *                 overlapped with the kernels), runs the same kernels, copies the results back and
 *                 synchronises before returning.
 *   - return value: 0 on success, otherwise a negative AMPSM_E* code or a positive cudaError_t;
 *     ampsm_last_error() returns a message for the calling thread.
 */
#ifndef AMPSM_B200_H
#define AMPSM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMPSM_MAX_K 16          /* largest alphabet of the reference (config.py:78-115) */
#define AMPSM_NUM_COUNTERS 24   /* 64-bit words in a counter block, see below */

#define AMPSM_EINVAL   (-1)     /* bad argument (dimension, NULL pointer, unsupported combination) */
#define AMPSM_ENOFIT   (-2)     /* problem does not fit the requested kernel */
#define AMPSM_ENODEV   (-3)     /* no sm_100 device */

/* Constellation exactly as config.symbols (complex128, config.py:117) and config.gray. */
typedef struct {
    int32_t K;
    int32_t gray[AMPSM_MAX_K];
    double  re[AMPSM_MAX_K];
    double  im[AMPSM_MAX_K];
} ampsm_alphabet;

/* Problem geometry and behaviour switches shared by all detectors. */
typedef struct {
    int32_t n;               /* rows of H / A:  Nr*Lout                                  (config.py:137) */
    int32_t N;               /* columns:        Nt*Lin                                                    */
    int32_t R;               /* VAMP only: number of singular values, min(n, N)         (vamp.py:28)      */
    int32_t Nt, Na, Nr;      /* antennas; section size M = Nt/Na, sections per frame L = Na*Lin            */
    int32_t Lin, Lout;       /* time slots in / out (SCAMP: Lc, Lr)                     (config.py:62-66) */
    int32_t max_iters;       /* config.N_Layers                                          (config.py:147)  */
    int32_t early_exit;      /* 1: per-frame torch.allclose exit (bamp.py:140), 0: exactly max_iters       */
    int32_t shift_mode;      /* 0: per-section max shift (finite everywhere);
                                1: frame-global max|x| in float64 as bamp.py:70 (NaN-faithful, slower)     */
    int32_t exp_f64;         /* 1: denoiser exponents/exp in float64 as the reference; 0: float32 exp      */
    int32_t decision;        /* 0: MAP decision (loss.py:282-302, mode 'sparc'); 1: segmented (223-250);
                                2: generator_mode 'random' -- i.i.d.-prior denoiser (bamp.py:79-97) and the top-Na
                                decision (loss.py:252-280); labels are then Lin*Na per frame; BAMP generic only */
    int32_t index_bits_kept; /* low bits of the index XOR that Loss.de2bi keeps (loss.py:20,168)           */
    int32_t kernel;          /* 0: auto, 1: generic shared-memory kernel, 2: register-resident kernel, one warp
                                per frame, 3: register-resident kernel, two warps per frame (64 x 32 shapes)   */
    int32_t reserved0;
    int64_t frame_base;      /* global index of the first frame of this call (flat indices, loss.py:300)   */
} ampsm_problem;

/*
 * Counter block (AMPSM_NUM_COUNTERS 64-bit words), ACCUMULATED into by every call (zero it first):
 *   [0] frames          [1] frames with any wrong entry (loss.py:150)
 *   [2] wrong time slots (loss.py:133)   [3] slot 0   [4] slot Lin/2   [5] last slot (loss.py:134-136)
 *   [6] wrong indices (loss.py:165)      [7] wrong Gray labels (loss.py:166)
 *   [8] index bit errors (loss.py:168)   [9] symbol bit errors (loss.py:172)
 *   [10] executed iterations summed over frames   [11] frames whose estimate holds a NaN
 *   [12..15] reserved
 *   [16] sum |xmmse - x|^2 (double)  [17] slot 0  [18] slot Lin/2  [19] last slot (loss.py:116-119)
 *   [20..23] reserved
 */

/* Library / device info. */
const char* ampsm_version(void);
const char* ampsm_last_error(void);
int ampsm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes);

/*
 * BAMP -- replaces bamp.BAMP.forward (bamp.py:116-143): Tracker init (12-25), up to max_iters BAMPLayer
 * iterations (59-64) with the section-wise denoiser (66-77), the allclose exit (140) and, when x_true is
 * given, Loss on (xmap, xmmse) (142; loss.py:67-179).
 *   H      : complex64 [n][N] shared by all frames (H_frame_stride = 0) or [frames][n][N]
 *            (H_frame_stride = n*N, in complex elements)
 *   y      : complex64 [frames][n]
 *   sigma2 : noise variance (Na/Nr)/SNR (bamp.py:111,134); sigma2_per_frame (device float[frames]) overrides
 *            the scalar when not NULL
 *   x_true : complex64 [frames][N], sym_true/idx_true : int64 [frames][L] Gray labels / flat non-zero
 *            positions as returned by Data.generate_message (data.py:88-90); all three NULL => no Loss
 *   xmap, xmmse : complex64 [frames][N];  var : float [frames][N];  iters : int32 [frames]
 *   traj   : float [frames][max_iters][3] = {mean tau, mean var, mean |xmmse-x|^2} per executed iteration
 *   counters : uint64 [AMPSM_NUM_COUNTERS] (see above)
 * Alignment: the register-resident kernels (p->kernel = 0 'auto' or 2 'fast') need H, y and x_true 16-byte aligned (as every
 * cudaMalloc / torch allocation is); with kernel = 0 anything else silently takes the generic kernel, with kernel = 2 the call
 * returns AMPSM_ENOFIT.
 */
int ampsm_bamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                      const void* H, int64_t H_frame_stride, const void* y,
                      double sigma2, const float* sigma2_per_frame,
                      const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                      void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                      uint64_t* counters, void* stream);

/*
 * BAMP on a structured ISI channel -- same detector as ampsm_bamp_detect (bamp.py:116-143), with the reference's
 * block-Toeplitz matrix (channel.py:53-72 generate_channel, 85-91 generate_as_sparc) given by its taps instead of the
 * dense (Nr*Lout) x (Nt*Lin) array: block (i, j) of H (output slot i, input slot j) is taps[i - j] for 0 <= i - j < Lh
 * and zero otherwise; with cyclic = 1 the difference is taken modulo Lin (channel_truncation 'cyclic').  'trunc' is
 * Lout = Lin, 'tail' is Lout = Lin + Lh - 1, both with cyclic = 0.
 *   taps : complex64 [Lh][Nr][Nt] shared by all frames (taps_frame_stride = 0) or [frames][Lh][Nr][Nt]
 *          (stride in complex elements); the scaling of channel.py:55 / 87-91 (sqrt(W) h) already applied.
 * Every other argument as ampsm_bamp_detect.  H, H^H, |H|^2 and |H|^2^T are applied as block convolutions from the taps
 * held in shared memory: Lh*Nr*Nt*8 bytes per frame instead of n*N*8 (24 x 128 x 3 taps: 72 KiB vs 25 MiB at Lin = 32).
 */
int ampsm_bamp_detect_taps(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                           const void* taps, int64_t taps_frame_stride, int32_t Lh, int32_t cyclic, const void* y,
                           double sigma2, const float* sigma2_per_frame,
                           const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                           void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                           uint64_t* counters, void* stream);

int ampsm_bamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                           const void* H, int64_t H_frame_stride, const void* y,
                           double sigma2, const float* sigma2_per_frame,
                           const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                           void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                           uint64_t* counters, int device);

/*
 * VAMP -- replaces vamp.VAMP.forward (vamp.py:159-191): Tracker (12-28), VAMPLayer iterations (66-94) with
 * the un-halved scalar-variance denoiser (96-119), exit on var (185), Loss on (r, xmmse) (187).
 *   U [n][R], s [R], Vh [R][N]: the caller's thin SVD (vamp_model.py:58), shared (stride 0) or per frame
 *   (strides in elements).  is_double = 1: U, Vh, y are complex128 and s is float64 (the reference fed with
 *   upcast inputs); xmap is then complex128, xmmse/var stay complex64/float32 (vamp.py:119).
 *   sparsity = Na/Nt (vamp.py:25-26).  traj : float [frames][max_iters][3] = {sigma2_tilde, mean var, mse}.
 */
int ampsm_vamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, int is_double,
                      const void* U, int64_t U_frame_stride, const void* s, int64_t s_frame_stride,
                      const void* Vh, int64_t Vh_frame_stride, const void* y,
                      double sigma2, const float* sigma2_per_frame, double sparsity,
                      const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                      void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                      uint64_t* counters, void* stream);

/*
 * The reference's second VAMP -- replaces vamp2.VAMP.forward (vamp2.py:104-131: "direct implementation of Rangan (with
 * damping)", imported by none of the reference's drivers): Tracker (12-26), VAMPLayer (52-76) with its own denoiser (78-87:
 * variance E|s|^2 - |E s|^2), exit on var (124), Loss on (r, xmmse) (128).  complex64 factors as ampsm_vamp_detect;
 * damping = the `damping` argument of vamp2.VAMP.  traj : float [frames][max_iters][3] = {gamma, mean var, mse}.
 */
int ampsm_vamp2_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                       const void* U, int64_t U_frame_stride, const void* s, int64_t s_frame_stride,
                       const void* Vh, int64_t Vh_frame_stride, const void* y,
                       double sigma2, const float* sigma2_per_frame, double damping,
                       const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                       void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                       uint64_t* counters, void* stream);

int ampsm_vamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, int is_double,
                           const void* U, int64_t U_frame_stride, const void* s, int64_t s_frame_stride,
                           const void* Vh, int64_t Vh_frame_stride, const void* y,
                           double sigma2, const float* sigma2_per_frame, double sparsity,
                           const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                           void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                           uint64_t* counters, int device);

/*
 * Batched thin SVD -- replaces the caller-side `U, s, Vh = torch.linalg.svd(A, full_matrices=False)` of the reference's
 * VAMP driver (vamp_model.py:56-58) for per-frame channel matrices, by one-sided Jacobi on the device (one warp per
 * matrix, csrc/svd_jacobi.cu).
 *   H  : complex64 [frames][n][N] dense, 1 <= n <= 32, n <= N, N in {8, 16, 32, 64}
 *   U  : complex64 [frames][n][n];  s : float [frames][n] descending;  Vh : complex64 [frames][n][N];  H = U diag(s) Vh
 *   sweeps : int32 [frames] executed Jacobi sweeps (optional, may be NULL)
 * The phases of the singular-vector pairs are not LAPACK's (VAMP is invariant to them: vamp.py:22,67,72).
 */
int ampsm_svd_batched(int64_t frames, int32_t n, int32_t N, const void* H, void* U, float* s, void* Vh,
                      int32_t* sweeps, void* stream);

/*
 * VAMP straight from per-frame channel matrices: ampsm_svd_batched into `workspace`, then ampsm_vamp_detect with
 * per-frame factors -- the whole of vamp_model.py:56-61 for one batch of frames.  complex64 only.
 *   H : complex64 [frames][n][N];  workspace : device scratch of ampsm_vamp_from_h_workspace_bytes(p, frames) bytes.
 *   Every other argument as ampsm_vamp_detect.  p->R must be min(n, N) = n.
 */
int64_t ampsm_vamp_from_h_workspace_bytes(const ampsm_problem* p, int64_t frames);
int ampsm_vamp_detect_from_h(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* H, const void* y,
                             double sigma2, const float* sigma2_per_frame, double sparsity,
                             const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                             void* xmap, void* xmmse, float* var, int32_t* iters, float* traj,
                             uint64_t* counters, void* workspace, void* stream);

/*
 * On-device generation of Monte-Carlo frames with a counter-based generator (Philox4x32-10,
 * csrc/framegen.cuh) -- the throughput-sweep replacement of the reference's per-epoch draws:
 *   channel.py:53-55   H = (N(0,1) + j N(0,1)) sqrt(h_var / 2), h_var = 1 / Nr
 *   data.py:74-91      one active antenna per section and one constellation point each -> x, Gray labels, flat positions
 *   channel.py:113-115 y = H x + (N(0,1) + j N(0,1)) sqrt(sigma2 / 2)
 * and BASELINE config 5's Kronecker correlation: H = Rr_root G Rt_root when the roots are given, or the exponential model
 * from rho_r / rho_t alone.  Frame f of the call is
 * frame counter_base + f of the stream `seed`: shards and chunks that cover the same global range draw the same frames.  The
 * draws are not the reference's numpy / torch sequences; parity subsets keep the reference's own RNG path.
 * Shapes: Lin = 1, 1 <= n <= 32, n <= N, N in {8, 16, 32, 64}, at most 32 sections, decision = 0.
 */
typedef struct {
    uint64_t seed;           /* Philox key */
    int64_t  counter_base;   /* global number of frame 0 of this call */
    double   h_var;          /* variance of a complex channel entry, 1 / Nr in the reference (channel.py:55) */
    const void* Rr_root;     /* complex64 [n][n] device pointer or NULL */
    const void* Rt_root;     /* complex64 [N][N] device pointer or NULL */
    int32_t  real_roots;     /* 1: the imaginary parts of both roots are zero (exponential correlation): half the work */
    int32_t  reserved0;
    double   rho_r, rho_t;   /* exponential correlation R[i][j] = rho^|i-j| on the receive / transmit side, generated by AR(1)
                                recursions (same distribution as the Hermitian roots, ~60x cheaper); 0 = none; not together with
                                the root of the same side */
} ampsm_gen;

/* H : complex64 [frames][n][N] and y : complex64 [frames][n] (either may be NULL); x : complex64 [frames][N],
 * sym / idx : int64 [frames][L] as Data.generate_message returns them (data.py:88-90; idx counts from p->frame_base). */
int ampsm_generate_frames(const ampsm_problem* p, const ampsm_alphabet* a, const ampsm_gen* gen, int64_t frames, double sigma2,
                          void* H, void* y, void* x, int64_t* sym, int64_t* idx, void* stream);

/*
 * VAMP on generated frames, the channel matrix never in HBM: the Jacobi SVD kernel draws every frame straight into its
 * shared-memory tile (the same stream as ampsm_generate_frames, bit for bit), factorises it and hands s, Vh and U^H y to
 * ampsm_vamp_detect -- vamp_model.py:44-61 for one batch of frames in two kernels.  x / sym / idx receive the ground truth the
 * Loss counts against (required); workspace as ampsm_vamp_from_h_workspace_bytes(p, frames).
 */
int ampsm_vamp_detect_generated(const ampsm_problem* p, const ampsm_alphabet* a, const ampsm_gen* gen, int64_t frames,
                                double sigma2, double sparsity, void* x, int64_t* sym, int64_t* idx,
                                void* xmap, void* xmmse, float* var, int32_t* iters, uint64_t* counters,
                                void* workspace, void* stream);

/*
 * SCAMP -- replaces scamp.SCAMP.forward (scamp.py:77-108): Tracker (8-25), SCAMPLayer iterations (43-59) with
 * the mean-only denoiser (61-68), exit on psi (105), Loss on (xmap, xmmse) (107).
 *   W : float [Lout][Lin] base matrix; A : complex64 [n][N] design matrix shared by all frames of the call
 *   workspace : device scratch of ampsm_scamp_workspace_bytes(p, frames) bytes (NULL: the library allocates
 *   and frees on the stream).  psi : float [frames][Lin] optional output.
 *   traj : float [frames][max_iters][3] = {mean tau, mean psi, mse}.
 */
int64_t ampsm_scamp_workspace_bytes(const ampsm_problem* p, int64_t frames);
int ampsm_scamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                       const float* W, const void* A, const void* y,
                       double sigma2, const float* sigma2_per_frame,
                       const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                       void* xmap, void* xmmse, float* psi, int32_t* iters, float* traj,
                       uint64_t* counters, void* workspace, void* stream);

/*
 * SCAMP on a STRUCTURED design matrix -- what Channel.generate_as_sparc builds (channel.py:76-96): block (r, c) of A is
 * taps[r - c] (Nr x Nt) for 0 <= r - c < Lh and zero elsewhere.  taps : complex64 [Lh][Nr][Nt]; the dense A (142 MB at
 * BASELINE config 4) is never formed: both mat-vecs run as dense tensor-core GEMMs over (frame, column block) rows fed by
 * tensor TMA (csrc/scamp_st.cu).  ampsm_scamp_taps_workspace_bytes returns AMPSM_ENOFIT when the shape has no structured
 * path (needs Lin <= 128, Lh * Nr <= 128, Nr even): run ampsm_scamp_detect on the dense matrix then.
 */
int64_t ampsm_scamp_taps_workspace_bytes(const ampsm_problem* p, int64_t frames, int32_t Lh);
int ampsm_scamp_detect_taps(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                            const float* W, const void* taps, int32_t Lh, const void* y,
                            double sigma2, const float* sigma2_per_frame,
                            const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                            void* xmap, void* xmmse, float* psi, int32_t* iters, float* traj,
                            uint64_t* counters, void* workspace, void* stream);

int ampsm_scamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                            const float* W, const void* A, const void* y,
                            double sigma2, const float* sigma2_per_frame,
                            const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                            void* xmap, void* xmmse, float* psi, int32_t* iters, float* traj,
                            uint64_t* counters, int device);

/*
 * Loss -- replaces loss.Loss.error_rate (loss.py:67-103) for estimates already on the device: hard decision
 * (MAP 282-302 or segmented 223-250) and the counters above.  iters may be NULL.
 */
int ampsm_loss_count(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames,
                     const void* xmap, const void* xmmse, const void* x_true,
                     const int64_t* sym_true, const int64_t* idx_true, const int32_t* iters,
                     uint64_t* counters, void* stream);

/*
 * Shrink -- replaces shrink.Shrink.forward / sw_shrinkOOK (shrink.py:45-75, 77-95, 139-157), the element-wise
 * denoisers of the reference's `random`-mode VAMP variant (vamp2.py:46), complex64 / float32 as the reference.
 *   kind  : AMPSM_SHRINK_BAYES  out_c[i] = Ps sum_k s_k G_k / (P0 G_0 + Ps sum_k G_k)           (shrink.py:77-95)
 *           AMPSM_SHRINK_OOK    out_f[i] = 1 / (1 + eta + 1e-9); *der_sum += sum_i nan_to_num(2 eta out^2 / cov),
 *                               the caller forms dxdr = der_sum / elems (zero *der_sum first)     (shrink.py:139-157)
 *           AMPSM_SHRINK_SW_OOK per section of M entries: out_c = Exp, out_f = Exp (1 - Exp)      (shrink.py:58-75)
 *   P0, Ps: config.P0 / config.Ps (config.py:76-114);  r : complex64 [elems];  cov : float [elems] (cov_stride = 1)
 *           or one value (cov_stride = 0, the 0-dim gamma of vamp2.py:59).
 * `shrink` and `lasso` raise in the reference itself (shrink.py:113, 135) and have no entry here.
 */
#define AMPSM_SHRINK_BAYES  0
#define AMPSM_SHRINK_OOK    1
#define AMPSM_SHRINK_SW_OOK 2
int ampsm_shrink(int kind, const ampsm_alphabet* a, double P0, double Ps, int64_t elems, int32_t M, const void* r,
                 const float* cov, int64_t cov_stride, void* out_c, float* out_f, double* der_sum, void* stream);

/*
 * NUMA placement of the host entry points (*_detect_host).  They pin the calling thread, for the duration of the call, to the
 * CPUs next to the GPU (sysfs local_cpulist of its PCI device; AMPSM_NO_NUMA_BIND=1 disables it).  ampsm_host_alloc returns
 * pinned host memory first-touched from such a thread, so that its pages live on the GPU's node: hand those buffers to the
 * host entry points.  ampsm_host_numa_info: 1 when binding narrows the thread's CPU set (a multi-node machine), else 0.
 */
int ampsm_host_alloc(int device, size_t bytes, void** out);
int ampsm_host_free(void* p);
int ampsm_host_numa_info(int device, int* n_local_cpus, int* n_allowed_cpus);

/*
 * Measurement helpers (bench.py): FP32 FFMA throughput of the device in TFLOP/s (the roofline denominator for
 * the shared-memory / register resident iterations, which MEASURED_PEAKS.json does not hold), and the number of
 * kernel launches this library has made since the last reset (bench.py's gpu_launches).
 */
int ampsm_probe_fp32_tflops(int device, double* tflops);
int ampsm_probe_fp32x2_tflops(int device, double* tflops);   /* same, issued as packed FFMA2 */
int ampsm_probe_fp64_tflops(int device, double* tflops);     /* DFMA: the denominator of the complex128 kernels */
int64_t ampsm_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* AMPSM_B200_H */
