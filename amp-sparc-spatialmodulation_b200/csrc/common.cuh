// Shared device/host helpers for the ampsm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/ampsm_b200.h"

#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ < 1000
#error "ampsm_b200 kernels are written for sm_100a (B200) only"
#endif

namespace ampsm {

constexpr int kWarp = 32;
constexpr float kRtol = 1e-5f;   // torch.allclose defaults (bamp.py:140)
constexpr float kAtol = 1e-8f;

// counter block layout (include/ampsm_b200.h)
enum Counter : int {
    C_FRAMES = 0, C_FRAME_ERR, C_SLOT_ERR, C_SLOT_FIRST, C_SLOT_MID, C_SLOT_LAST, C_INDEX_ERR, C_SYMBOL_ERR,
    C_INDEX_BIT, C_SYMBOL_BIT, C_ITERS, C_NAN_FRAMES, C_NUM_INT = 12, C_SQERR = 16, C_SQ_FIRST, C_SQ_MID, C_SQ_LAST
};

// Alphabet in kernel-parameter (constant bank) space: float64 as the reference's config.symbols plus the
// float32 roundings used by the fast exponent path and by the complex64 value compare of Loss.
struct DevAlphabet {
    int K;
    int sbits;
    int gray[AMPSM_MAX_K];
    double re[AMPSM_MAX_K], im[AMPSM_MAX_K];
    float ref[AMPSM_MAX_K], imf[AMPSM_MAX_K];
    float rel[AMPSM_MAX_K], iml[AMPSM_MAX_K];   // float32 residuals re - ref, im - imf (compensated float32 exponents)
};

// Geometry handed to every kernel.
struct Geom {
    int n, N, R;
    int Nt, Na, Nr, Lin, Lout;
    int M, L;            // section size, sections per frame
    int max_iters;
    int early_exit, shift_mode, decision, index_bits_kept;
    long long frame_base;
};

struct LossIO {
    const float2* x_true;        // [frames][N] or nullptr
    const long long* sym_true;   // [frames][L]
    const long long* idx_true;   // [frames][L]
    unsigned long long* counters;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
// max that lets NaN win (np.max / torch.max semantics, bamp.py:70)
__device__ __forceinline__ double nanmax(double a, double b) { return (a != a) ? a : ((b != b) ? b : (a > b ? a : b)); }
__device__ __forceinline__ double warp_nanmax(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = nanmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier + 1-D bulk TMA (cp.async.bulk), the staging path for per-frame matrices -------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// complex64 / real as c10::complex and numpy evaluate it for a real divisor: multiply by the rounded reciprocal
__device__ __forceinline__ float2 cdiv_real(float2 a, float d) {
    float r = __frcp_rn(d);
    return make_float2(__fmul_rn(a.x, r), __fmul_rn(a.y, r));
}

}  // namespace ampsm

// host-side helpers -----------------------------------------------------------------------------------------------
namespace ampsm {
void set_error(const char* fmt, ...);
void count_launch();
int check_cuda(cudaError_t e, const char* what);
}  // namespace ampsm
