// C-ABI entry points of libampsm_b200.so (declared in include/ampsm_b200.h): argument checking, kernel selection,
// and the host-buffer variants that overlap chunked host<->device copies with the kernels on two streams.
#include <cstdlib>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <functional>
#include <mutex>
#include <vector>

#include <sched.h>
#include <unistd.h>

#include "kernels.h"

namespace ampsm {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int check_cuda(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (int)e;
}

static int make_geom(const ampsm_problem* p, const ampsm_alphabet* a, Geom* g, DevAlphabet* al, bool need_R) {
    if (!p || !a) { set_error("problem / alphabet pointer is NULL"); return AMPSM_EINVAL; }
    if (a->K < 1 || a->K > AMPSM_MAX_K) { set_error("alphabet size K=%d outside 1..%d", a->K, AMPSM_MAX_K); return AMPSM_EINVAL; }
    if (p->n < 1 || p->N < 1 || p->Nt < 1 || p->Na < 1 || p->Nr < 1 || p->Lin < 1 || p->Lout < 1 || p->max_iters < 1) {
        set_error("non-positive dimension in ampsm_problem"); return AMPSM_EINVAL;
    }
    const bool iid = p->decision == 2;          // generator_mode 'random': i.i.d. prior, Na active entries anywhere in a time slot
    if (p->decision < 0 || p->decision > 2) { set_error("decision=%d outside 0..2", p->decision); return AMPSM_EINVAL; }
    if (!iid && p->Nt % p->Na != 0) { set_error("Na=%d must divide Nt=%d (sectioned modes)", p->Na, p->Nt); return AMPSM_EINVAL; }
    if (iid && (p->Na > p->Nt || p->Na > 32 || p->Nt > 1024)) { set_error("random mode needs Na <= min(Nt, 32) and Nt <= 1024"); return AMPSM_EINVAL; }
    if (p->N != p->Nt * p->Lin || p->n != p->Nr * p->Lout) {
        set_error("N must equal Nt*Lin and n must equal Nr*Lout (got N=%d n=%d)", p->N, p->n); return AMPSM_EINVAL;
    }
    if (need_R && (p->R < 1 || p->R > p->N)) { set_error("R=%d outside 1..N", p->R); return AMPSM_EINVAL; }
    if (p->shift_mode == 1 && !p->exp_f64) { set_error("shift_mode=1 (reference shift) needs exp_f64=1"); return AMPSM_EINVAL; }
    if (p->index_bits_kept < 0 || p->index_bits_kept > 64) { set_error("index_bits_kept outside 0..64"); return AMPSM_EINVAL; }
    g->n = p->n; g->N = p->N; g->R = p->R;
    g->Nt = p->Nt; g->Na = p->Na; g->Nr = p->Nr; g->Lin = p->Lin; g->Lout = p->Lout;
    g->M = iid ? p->Nt : p->Nt / p->Na;          // random mode: a 'section' is one time slot of Nt entries holding Na labels
    g->L = iid ? p->Lin : p->Na * p->Lin;
    g->max_iters = p->max_iters; g->early_exit = p->early_exit; g->shift_mode = p->shift_mode;
    g->decision = p->decision; g->index_bits_kept = p->index_bits_kept; g->frame_base = p->frame_base;
    al->K = a->K;
    int sb = 0;
    while ((1 << (sb + 1)) <= a->K) ++sb;            // int(log2(K)), config.py:119
    al->sbits = sb;
    for (int k = 0; k < AMPSM_MAX_K; ++k) {
        const bool in = k < a->K;
        al->gray[k] = in ? a->gray[k] : 0;
        al->re[k] = in ? a->re[k] : 0.0;
        al->im[k] = in ? a->im[k] : 0.0;
        al->ref[k] = (float)al->re[k];
        al->imf[k] = (float)al->im[k];
        al->rel[k] = (float)(al->re[k] - (double)al->ref[k]);
        al->iml[k] = (float)(al->im[k] - (double)al->imf[k]);
    }
    return 0;
}

static int check_loss_io(const void* x_true, const int64_t* sym, const int64_t* idx) {
    const int have = (x_true != nullptr) + (sym != nullptr) + (idx != nullptr);
    if (have != 0 && have != 3) { set_error("x_true, sym_true and idx_true must be given together"); return AMPSM_EINVAL; }
    return 0;
}

// ---- NUMA placement of the host path ----------------------------------------------------------------------------
// The copies of the host entry points are PCIe transfers from the caller's buffers: on a two-socket box a buffer (or the
// thread issuing the copies) on the other socket crosses the inter-socket link (SCALE_r01: 22 GB/s per GPU at 8 GPUs against
// 55 GB/s alone).  The CPUs next to a GPU are read from sysfs (local_cpulist of its PCI device); HostBind pins the calling
// thread to them for its lifetime, ampsm_host_alloc first-touches pinned memory from such a thread so that its pages live on
// the GPU's node.  On a single-node machine (or when sysfs says nothing) both are no-ops.
static bool gpu_local_cpus(int device, cpu_set_t* set) {
    char bus[32] = "";
    if (cudaDeviceGetPCIBusId(bus, sizeof(bus), device) != cudaSuccess) return false;
    for (char* c = bus; *c; ++c) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/local_cpulist", bus);
    FILE* f = fopen(path, "r");
    if (!f) return false;
    char line[1024] = "";
    const bool got = fgets(line, sizeof(line), f) != nullptr;
    fclose(f);
    if (!got) return false;
    CPU_ZERO(set);
    int count = 0;
    for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {      // "0-15,32-47"
        int lo = 0, hi = 0;
        const int k = sscanf(tok, "%d-%d", &lo, &hi);
        if (k < 1) continue;
        if (k == 1) hi = lo;
        for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c) { CPU_SET(c, set); ++count; }
    }
    return count > 0;
}
struct HostBind {
    cpu_set_t old;
    bool bound = false;
    explicit HostBind(int device) {
        if (getenv("AMPSM_NO_NUMA_BIND")) return;
        cpu_set_t local, want;
        if (sched_getaffinity(0, sizeof(old), &old) != 0 || !gpu_local_cpus(device, &local)) return;
        CPU_AND(&want, &old, &local);
        if (CPU_COUNT(&want) == 0 || CPU_EQUAL(&want, &old)) return;      // nothing to narrow (single node, or an outer binding)
        bound = sched_setaffinity(0, sizeof(want), &want) == 0;
    }
    ~HostBind() {
        if (bound) sched_setaffinity(0, sizeof(old), &old);
    }
};

// ---- host-buffer runner: chunks of frames, two streams, grow-only per-device arenas ---------------------------
struct Arena {
    unsigned char* base = nullptr;
    size_t cap = 0;
    size_t used = 0;
    int reserve(size_t bytes) {
        if (bytes <= cap) return 0;
        if (base) cudaFree(base);
        base = nullptr; cap = 0;
        if (int e = check_cuda(cudaMalloc((void**)&base, bytes), "cudaMalloc(host-path arena)")) return e;
        cap = bytes;
        return 0;
    }
    void* take(size_t bytes) {
        unsigned char* p = base + used;
        used += (bytes + 255) & ~size_t(255);
        return p;
    }
};
struct DeviceCtx {
    std::mutex mu;
    cudaStream_t stream[2] = {nullptr, nullptr};
    Arena slot[2], shared;
};
static DeviceCtx* device_ctx(int device) {
    static std::mutex mu;
    static std::vector<DeviceCtx*> ctx;
    std::lock_guard<std::mutex> lk(mu);
    if ((int)ctx.size() <= device) ctx.resize(device + 1, nullptr);
    if (!ctx[device]) ctx[device] = new DeviceCtx();
    return ctx[device];
}

struct Field {
    const void* h_in;
    void* h_out;
    size_t bpf;      // bytes per frame
};
using LaunchFn = std::function<int(long long f0, long long nf, void* const* d, unsigned long long* d_counters,
                                   void* scratch, cudaStream_t st)>;

static int run_host(int device, long long frames, const std::vector<Field>& fields,
                    const std::vector<std::pair<const void*, size_t>>& shared_in, std::vector<void*>* shared_dev,
                    size_t scratch_per_frame, size_t scratch_fixed, uint64_t* h_counters, const LaunchFn& launch) {
    if (int e = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return e;
    HostBind bind(device);                       // the copies are issued from a CPU next to the GPU
    DeviceCtx* cx = device_ctx(device);
    std::lock_guard<std::mutex> lk(cx->mu);
    for (int s = 0; s < 2; ++s)
        if (!cx->stream[s])
            if (int e = check_cuda(cudaStreamCreateWithFlags(&cx->stream[s], cudaStreamNonBlocking), "cudaStreamCreate")) return e;
    size_t per_frame = scratch_per_frame;
    for (const Field& f : fields) per_frame += ((f.h_in || f.h_out) ? f.bpf : 0);
    if (per_frame == 0) per_frame = 1;
    const size_t target = (size_t)96 << 20;                        // ~96 MiB of traffic per chunk
    long long chunk = (long long)(target / per_frame);
    if (chunk < 1) chunk = 1;
    if (chunk > frames) chunk = frames;
    // shared inputs + counters
    size_t shared_bytes = 256 + AMPSM_NUM_COUNTERS * 8;
    for (auto& s : shared_in) shared_bytes += (s.second + 255) & ~size_t(255);
    if (int e = cx->shared.reserve(shared_bytes)) return e;
    cx->shared.used = 0;
    unsigned long long* d_counters = (unsigned long long*)cx->shared.take(AMPSM_NUM_COUNTERS * 8);
    if (int e = check_cuda(cudaMemsetAsync(d_counters, 0, AMPSM_NUM_COUNTERS * 8, cx->stream[0]), "memset counters")) return e;
    shared_dev->clear();
    for (auto& s : shared_in) {
        void* d = s.first ? cx->shared.take(s.second) : nullptr;
        if (d)
            if (int e = check_cuda(cudaMemcpyAsync(d, s.first, s.second, cudaMemcpyHostToDevice, cx->stream[0]), "H2D shared")) return e;
        shared_dev->push_back(d);
    }
    if (int e = check_cuda(cudaStreamSynchronize(cx->stream[0]), "sync shared inputs")) return e;
    size_t slot_bytes = scratch_fixed + 256 + (size_t)chunk * scratch_per_frame + 256;
    for (const Field& f : fields) slot_bytes += (((size_t)chunk * f.bpf) + 255) & ~size_t(255);
    for (int s = 0; s < 2; ++s)
        if (int e = cx->slot[s].reserve(slot_bytes)) return e;
    int rc = 0;
    int which = 0;
    for (long long f0 = 0; f0 < frames && rc == 0; f0 += chunk, which ^= 1) {
        const long long nf = (frames - f0 < chunk) ? frames - f0 : chunk;
        Arena& ar = cx->slot[which];
        cudaStream_t st = cx->stream[which];
        ar.used = 0;                                              // stream order protects the previous use of this slot
        std::vector<void*> d(fields.size(), nullptr);
        for (size_t i = 0; i < fields.size(); ++i) {
            const Field& f = fields[i];
            if (!f.h_in && !f.h_out) continue;
            d[i] = ar.take((size_t)chunk * f.bpf);
            if (f.h_in)
                rc = check_cuda(cudaMemcpyAsync(d[i], (const unsigned char*)f.h_in + (size_t)f0 * f.bpf, (size_t)nf * f.bpf,
                                                cudaMemcpyHostToDevice, st), "H2D chunk");
            if (rc) break;
        }
        if (rc) break;
        void* scratch = (scratch_per_frame || scratch_fixed) ? ar.take(scratch_fixed + (size_t)chunk * scratch_per_frame) : nullptr;
        rc = launch(f0, nf, d.data(), d_counters, scratch, st);
        if (rc) break;
        for (size_t i = 0; i < fields.size(); ++i) {
            const Field& f = fields[i];
            if (!f.h_out) continue;
            rc = check_cuda(cudaMemcpyAsync((unsigned char*)f.h_out + (size_t)f0 * f.bpf, d[i], (size_t)nf * f.bpf,
                                            cudaMemcpyDeviceToHost, st), "D2H chunk");
            if (rc) break;
        }
    }
    for (int s = 0; s < 2; ++s) {
        const int e = check_cuda(cudaStreamSynchronize(cx->stream[s]), "sync host path");
        if (!rc) rc = e;
    }
    if (!rc && h_counters) {
        unsigned long long tmp[AMPSM_NUM_COUNTERS];
        rc = check_cuda(cudaMemcpy(tmp, d_counters, sizeof(tmp), cudaMemcpyDeviceToHost), "D2H counters");
        if (!rc) {
            for (int i = 0; i < 16; ++i) h_counters[i] += tmp[i];
            for (int i = 16; i < 20; ++i) {
                double a, b;
                memcpy(&a, &h_counters[i], 8);
                memcpy(&b, &tmp[i], 8);
                a += b;
                memcpy(&h_counters[i], &a, 8);
            }
        }
    }
    return rc;
}

}  // namespace ampsm

using namespace ampsm;

extern "C" {

const char* ampsm_version(void) { return "ampsm_b200 0.1 (sm_100a)"; }
const char* ampsm_last_error(void) { return g_err; }

int ampsm_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, int64_t* smem_optin_bytes) {
    cudaDeviceProp prop;
    if (int e = check_cuda(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return e;
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin_bytes) *smem_optin_bytes = (int64_t)prop.sharedMemPerBlockOptin;
    return 0;
}

int64_t ampsm_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

// Pinned host memory whose pages live on the NUMA node of `device`: allocated and first-touched by a thread bound to the CPUs
// next to that GPU.  For the buffers handed to the *_detect_host entry points.
int ampsm_host_alloc(int device, size_t bytes, void** out) {
    if (!out) { set_error("ampsm_host_alloc: out is NULL"); return AMPSM_EINVAL; }
    *out = nullptr;
    if (int e = check_cuda(cudaSetDevice(device), "cudaSetDevice")) return e;
    HostBind bind(device);
    void* p = nullptr;
    if (int e = check_cuda(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc")) return e;
    memset(p, 0, bytes);                         // first touch from the bound thread
    *out = p;
    return 0;
}
int ampsm_host_free(void* p) { return p ? check_cuda(cudaFreeHost(p), "cudaFreeHost") : 0; }
// 1 when the calling thread would be narrowed to CPUs next to `device` (a multi-node machine), 0 otherwise; n_local = their number
int ampsm_host_numa_info(int device, int* n_local, int* n_allowed) {
    cpu_set_t cur, local, want;
    if (n_local) *n_local = 0;
    if (n_allowed) *n_allowed = 0;
    if (sched_getaffinity(0, sizeof(cur), &cur) != 0) return 0;
    if (n_allowed) *n_allowed = CPU_COUNT(&cur);
    if (!gpu_local_cpus(device, &local)) return 0;
    CPU_AND(&want, &cur, &local);
    if (n_local) *n_local = CPU_COUNT(&want);
    return CPU_COUNT(&want) > 0 && !CPU_EQUAL(&want, &cur);
}

int ampsm_probe_fp32_tflops(int device, double* tflops) { return probe_fp32(device, tflops); }
int ampsm_probe_fp32x2_tflops(int device, double* tflops) { return probe_fp32x2(device, tflops); }
int ampsm_probe_fp64_tflops(int device, double* tflops) { return probe_fp64(device, tflops); }

// ---------------------------------------------------------------- BAMP
int ampsm_bamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* H,
                      int64_t H_frame_stride, const void* y, double sigma2, const float* sigma2_per_frame,
                      const void* x_true, const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse,
                      float* var, int32_t* iters, float* traj, uint64_t* counters, void* stream) {
    BampArgs k{};
    if (int e = make_geom(p, a, &k.g, &k.al, false)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || (frames > 0 && (!H || !y))) { set_error("BAMP: H / y is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (H_frame_stride != 0 && H_frame_stride < (int64_t)p->n * p->N) { set_error("BAMP: H_frame_stride smaller than n*N"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    if (p->decision == 2 && p->kernel > 1) { set_error("BAMP random mode runs the generic kernel only"); return AMPSM_ENOFIT; }
    k.H = (const float2*)H; k.H_stride = H_frame_stride; k.y = (const float2*)y;
    k.sigma2 = (float)sigma2; k.sigma2_pf = sigma2_per_frame;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = (float2*)xmap; k.xmmse = (float2*)xmmse; k.var = var; k.iters = iters; k.traj = traj; k.frames = frames;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->kernel != 1 && !p->exp_f64 && p->shift_mode == 0 && p->decision != 2) {
        const int rc = launch_bamp_fast(k, st);
        if (rc != AMPSM_ENOFIT) return rc;
        if (p->kernel == 2) return rc;
    } else if (p->kernel == 2) {
        set_error("BAMP register-resident kernel supports exp_f64=0, shift_mode=0 only");
        return AMPSM_ENOFIT;
    }
    return launch_bamp_generic(k, p->exp_f64 != 0, st);
}

int ampsm_bamp_detect_taps(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* taps,
                           int64_t taps_frame_stride, int32_t Lh, int32_t cyclic, const void* y, double sigma2,
                           const float* sigma2_per_frame, const void* x_true, const int64_t* sym_true, const int64_t* idx_true,
                           void* xmap, void* xmmse, float* var, int32_t* iters, float* traj, uint64_t* counters, void* stream) {
    BampArgs k{};
    if (int e = make_geom(p, a, &k.g, &k.al, false)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || (frames > 0 && (!taps || !y))) { set_error("BAMP taps: taps / y is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (Lh < 1 || (cyclic != 0 && cyclic != 1)) { set_error("BAMP taps: Lh < 1 or cyclic not in {0, 1}"); return AMPSM_EINVAL; }
    if (cyclic && (p->Lout != p->Lin || Lh > p->Lin)) { set_error("BAMP taps: cyclic operator needs Lout == Lin and Lh <= Lin"); return AMPSM_EINVAL; }
    if (!cyclic && (p->Lout < p->Lin || p->Lout > p->Lin + Lh - 1)) { set_error("BAMP taps: Lout=%d outside Lin..Lin+Lh-1", p->Lout); return AMPSM_EINVAL; }
    const int64_t tap_elems = (int64_t)Lh * p->Nr * p->Nt;
    if (taps_frame_stride != 0 && taps_frame_stride < tap_elems) { set_error("BAMP taps: taps_frame_stride smaller than Lh*Nr*Nt"); return AMPSM_EINVAL; }
    if (p->kernel > 1) { set_error("BAMP taps: the structured operator runs in the generic kernel only"); return AMPSM_ENOFIT; }
    if (frames == 0) return 0;
    k.taps = (const float2*)taps; k.taps_stride = taps_frame_stride; k.Lh = Lh; k.cyclic = cyclic;
    k.y = (const float2*)y; k.sigma2 = (float)sigma2; k.sigma2_pf = sigma2_per_frame;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = (float2*)xmap; k.xmmse = (float2*)xmmse; k.var = var; k.iters = iters; k.traj = traj; k.frames = frames;
    return launch_bamp_generic(k, p->exp_f64 != 0, (cudaStream_t)stream);
}

int ampsm_bamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* H,
                           int64_t H_frame_stride, const void* y, double sigma2, const float* sigma2_per_frame,
                           const void* x_true, const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse,
                           float* var, int32_t* iters, float* traj, uint64_t* counters, int device) {
    if (!p || !a) { set_error("problem / alphabet pointer is NULL"); return AMPSM_EINVAL; }
    if (frames <= 0) return frames < 0 ? AMPSM_EINVAL : 0;
    const size_t nN = (size_t)p->n * p->N, L = (size_t)p->Na * p->Lin;
    const bool per_frame_H = H_frame_stride != 0;
    if (per_frame_H && H_frame_stride != (int64_t)nN) { set_error("host path needs densely packed per-frame H"); return AMPSM_EINVAL; }
    std::vector<Field> f = {
        {per_frame_H ? H : nullptr, nullptr, nN * 8},          // 0
        {y, nullptr, (size_t)p->n * 8},                        // 1
        {sigma2_per_frame, nullptr, 4},                        // 2
        {x_true, nullptr, (size_t)p->N * 8},                   // 3
        {sym_true, nullptr, L * 8},                            // 4
        {idx_true, nullptr, L * 8},                            // 5
        {nullptr, xmap, (size_t)p->N * 8},                     // 6
        {nullptr, xmmse, (size_t)p->N * 8},                    // 7
        {nullptr, var, (size_t)p->N * 4},                      // 8
        {nullptr, iters, 4},                                   // 9
        {nullptr, traj, (size_t)p->max_iters * 12},            // 10
    };
    std::vector<std::pair<const void*, size_t>> shared = {{per_frame_H ? nullptr : H, nN * 8}};
    std::vector<void*> sd;
    return run_host(device, frames, f, shared, &sd, 0, 0, counters,
                    [&](long long f0, long long nf, void* const* d, unsigned long long* dc, void*, cudaStream_t st) {
                        ampsm_problem q = *p;
                        q.frame_base = p->frame_base + f0;
                        return ampsm_bamp_detect(&q, a, nf, per_frame_H ? d[0] : sd[0], per_frame_H ? (int64_t)nN : 0, d[1], sigma2,
                                                 (const float*)d[2], d[3], (const int64_t*)d[4], (const int64_t*)d[5], d[6], d[7],
                                                 (float*)d[8], (int32_t*)d[9], (float*)d[10], (uint64_t*)dc, st);
                    });
}

// ---------------------------------------------------------------- VAMP
int ampsm_vamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, int is_double, const void* U,
                      int64_t U_frame_stride, const void* s, int64_t s_frame_stride, const void* Vh, int64_t Vh_frame_stride,
                      const void* y, double sigma2, const float* sigma2_per_frame, double sparsity, const void* x_true,
                      const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse, float* var, int32_t* iters,
                      float* traj, uint64_t* counters, void* stream) {
    VampArgs k{};
    if (p && p->decision == 2) { set_error("VAMP has no working random mode in the reference (vamp.py:84 unpacks two values from a denoiser that returns one)"); return AMPSM_EINVAL; }
    if (int e = make_geom(p, a, &k.g, &k.al, true)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || (frames > 0 && (!U || !s || !Vh || !y))) { set_error("VAMP: U / s / Vh / y is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    k.U = U; k.U_stride = U_frame_stride; k.s = s; k.s_stride = s_frame_stride; k.Vh = Vh; k.Vh_stride = Vh_frame_stride;
    k.y = y; k.sigma2_d = sigma2; k.sigma2_pf = sigma2_per_frame; k.sparsity = sparsity;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = xmap; k.xmmse = (float2*)xmmse; k.var = var; k.iters = iters; k.traj = traj; k.frames = frames;
    if (p->kernel != 1 && !is_double && !p->exp_f64 && p->shift_mode == 0) {
        int rc = launch_vamp_fast(k, (cudaStream_t)stream);                         // 32 x 64 factors, one warp per frame
        if (rc == AMPSM_ENOFIT) rc = launch_vamp_quad(k, (cudaStream_t)stream);     // 64 x 128 factors, four warps per frame
        if (rc != AMPSM_ENOFIT || p->kernel == 2) return rc;
    } else if (p->kernel == 2) {
        set_error("VAMP register-resident kernel supports complex64, exp_f64=0, shift_mode=0 only");
        return AMPSM_ENOFIT;
    }
    if (is_double && p->kernel != 1) {                       // complex128 64 x 128 factors: register-resident DFMA kernel
        const int rc = launch_vamp_dbl(k, (cudaStream_t)stream);
        if (rc != AMPSM_ENOFIT) return rc;
    }
    return launch_vamp_generic(k, is_double != 0, p->exp_f64 != 0, (cudaStream_t)stream);
}

// vamp2.py: the damped direct form (csrc/vamp2.cu); complex64, generic kernel
int ampsm_vamp2_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* U, int64_t U_frame_stride,
                       const void* s, int64_t s_frame_stride, const void* Vh, int64_t Vh_frame_stride, const void* y, double sigma2,
                       const float* sigma2_per_frame, double damping, const void* x_true, const int64_t* sym_true,
                       const int64_t* idx_true, void* xmap, void* xmmse, float* var, int32_t* iters, float* traj, uint64_t* counters,
                       void* stream) {
    VampArgs k{};
    if (p && p->decision == 2) { set_error("vamp2: the Shrink('bayes') denoiser of mode 'random' is ampsm_shrink; the detector runs the sectioned modes"); return AMPSM_EINVAL; }
    if (int e = make_geom(p, a, &k.g, &k.al, true)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || (frames > 0 && (!U || !s || !Vh || !y))) { set_error("vamp2: U / s / Vh / y is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    k.U = U; k.U_stride = U_frame_stride; k.s = s; k.s_stride = s_frame_stride; k.Vh = Vh; k.Vh_stride = Vh_frame_stride;
    k.y = y; k.sigma2_d = sigma2; k.sigma2_pf = sigma2_per_frame; k.damping = (float)damping;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = xmap; k.xmmse = (float2*)xmmse; k.var = var; k.iters = iters; k.traj = traj; k.frames = frames;
    return launch_vamp2(k, p->exp_f64 != 0, (cudaStream_t)stream);
}

int ampsm_vamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, int is_double, const void* U,
                           int64_t U_frame_stride, const void* s, int64_t s_frame_stride, const void* Vh,
                           int64_t Vh_frame_stride, const void* y, double sigma2, const float* sigma2_per_frame,
                           double sparsity, const void* x_true, const int64_t* sym_true, const int64_t* idx_true, void* xmap,
                           void* xmmse, float* var, int32_t* iters, float* traj, uint64_t* counters, int device) {
    if (!p || !a) { set_error("problem / alphabet pointer is NULL"); return AMPSM_EINVAL; }
    if (frames <= 0) return frames < 0 ? AMPSM_EINVAL : 0;
    const size_t cs = is_double ? 16 : 8, rs = is_double ? 8 : 4;
    const size_t nR = (size_t)p->n * p->R, RN = (size_t)p->R * p->N, L = (size_t)p->Na * p->Lin;
    const bool pfU = U_frame_stride != 0, pfs = s_frame_stride != 0, pfV = Vh_frame_stride != 0;
    if ((pfU && U_frame_stride != (int64_t)nR) || (pfs && s_frame_stride != p->R) || (pfV && Vh_frame_stride != (int64_t)RN)) {
        set_error("host path needs densely packed per-frame factors"); return AMPSM_EINVAL;
    }
    std::vector<Field> f = {
        {pfU ? U : nullptr, nullptr, nR * cs},                 // 0
        {pfs ? s : nullptr, nullptr, (size_t)p->R * rs},       // 1
        {pfV ? Vh : nullptr, nullptr, RN * cs},                // 2
        {y, nullptr, (size_t)p->n * cs},                       // 3
        {sigma2_per_frame, nullptr, 4},                        // 4
        {x_true, nullptr, (size_t)p->N * 8},                   // 5
        {sym_true, nullptr, L * 8},                            // 6
        {idx_true, nullptr, L * 8},                            // 7
        {nullptr, xmap, (size_t)p->N * cs},                    // 8
        {nullptr, xmmse, (size_t)p->N * 8},                    // 9
        {nullptr, var, (size_t)p->N * 4},                      // 10
        {nullptr, iters, 4},                                   // 11
        {nullptr, traj, (size_t)p->max_iters * 12},            // 12
    };
    std::vector<std::pair<const void*, size_t>> shared = {
        {pfU ? nullptr : U, nR * cs}, {pfs ? nullptr : s, (size_t)p->R * rs}, {pfV ? nullptr : Vh, RN * cs}};
    std::vector<void*> sd;
    return run_host(device, frames, f, shared, &sd, 0, 0, counters,
                    [&](long long f0, long long nf, void* const* d, unsigned long long* dc, void*, cudaStream_t st) {
                        ampsm_problem q = *p;
                        q.frame_base = p->frame_base + f0;
                        return ampsm_vamp_detect(&q, a, nf, is_double, pfU ? d[0] : sd[0], pfU ? (int64_t)nR : 0, pfs ? d[1] : sd[1],
                                                 pfs ? p->R : 0, pfV ? d[2] : sd[2], pfV ? (int64_t)RN : 0, d[3], sigma2,
                                                 (const float*)d[4], sparsity, d[5], (const int64_t*)d[6], (const int64_t*)d[7], d[8],
                                                 d[9], (float*)d[10], (int32_t*)d[11], (float*)d[12], (uint64_t*)dc, st);
                    });
}

// ---------------------------------------------------------------- batched SVD, VAMP from the channel matrices
int ampsm_svd_batched(int64_t frames, int32_t n, int32_t N, const void* H, void* U, float* s, void* Vh, int32_t* sweeps,
                      void* stream) {
    if (frames < 0 || (frames > 0 && (!H || !U || !s || !Vh))) { set_error("SVD: NULL pointer or frames < 0"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    return launch_svd_jacobi((const float2*)H, frames, n, N, (float2*)U, s, (float2*)Vh, sweeps, nullptr, nullptr, (cudaStream_t)stream);
}

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// workspace: U^H y [frames][n] | s [frames][n] | Vh [frames][n][N] | one n x n identity (the `U` the detector is handed)
int64_t ampsm_vamp_from_h_workspace_bytes(const ampsm_problem* p, int64_t frames) {
    if (!p || frames < 0) return -1;
    const size_t n = p->n, N = p->N;
    return (int64_t)(align256((size_t)frames * n * 8) + align256((size_t)frames * n * 4) + align256((size_t)frames * n * N * 8) +
                     align256(n * n * 8));
}

int ampsm_vamp_detect_from_h(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* H, const void* y,
                             double sigma2, const float* sigma2_per_frame, double sparsity, const void* x_true,
                             const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse, float* var,
                             int32_t* iters, float* traj, uint64_t* counters, void* workspace, void* stream) {
    if (!p || !a) { set_error("problem / alphabet pointer is NULL"); return AMPSM_EINVAL; }
    if (frames < 0 || (frames > 0 && (!H || !y || !workspace))) { set_error("VAMP from H: H / y / workspace is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (p->R != p->n || p->n > p->N) { set_error("VAMP from H: needs R = n <= N (got n=%d N=%d R=%d)", p->n, p->N, p->R); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    const size_t n = p->n, N = p->N;
    unsigned char* w = (unsigned char*)workspace;
    float2* yrot = (float2*)w;
    float* s = (float*)(w + align256((size_t)frames * n * 8));
    void* Vh = w + align256((size_t)frames * n * 8) + align256((size_t)frames * n * 4);
    float2* eye = (float2*)((unsigned char*)Vh + align256((size_t)frames * n * N * 8));
    // VAMP needs U only in y~ = diag(s) U^H y (vamp.py:22): the Jacobi kernel rotates y along with the rows of H, and the
    // detector is handed U = I (shared by all frames) with U^H y in place of y
    if (int e = launch_svd_jacobi((const float2*)H, frames, p->n, p->N, nullptr, s, (float2*)Vh, nullptr, (const float2*)y, yrot,
                                  (cudaStream_t)stream))
        return e;
    if (int e = launch_identity(eye, p->n, (cudaStream_t)stream)) return e;
    return ampsm_vamp_detect(p, a, frames, 0, eye, 0, s, (int64_t)n, Vh, (int64_t)(n * N), yrot, sigma2, sigma2_per_frame,
                             sparsity, x_true, sym_true, idx_true, xmap, xmmse, var, iters, traj, counters, stream);
}

// ---------------------------------------------------------------- on-device frame generation (csrc/framegen.cuh)
static int make_gen(const ampsm_problem* p, const ampsm_alphabet* a, const ampsm_gen* gen, double sigma2, void* x, int64_t* sym, int64_t* idx,
                    Geom* g, GenArgs* ga) {
    DevAlphabet al{};
    if (int e = make_geom(p, a, g, &al, false)) return e;
    if (!gen) { set_error("frame generation: ampsm_gen pointer is NULL"); return AMPSM_EINVAL; }
    if (p->decision != 0) { set_error("frame generation draws sectioned messages (decision = 0; data.py:74-91)"); return AMPSM_EINVAL; }
    if (!(gen->h_var > 0.0) || !(sigma2 >= 0.0)) { set_error("frame generation: h_var must be positive and sigma2 non-negative"); return AMPSM_EINVAL; }
    if (check_loss_io(x, sym, idx)) return AMPSM_EINVAL;
    ga->seed = gen->seed;
    ga->counter_base = gen->counter_base;
    ga->h_std = (float)sqrt(gen->h_var / 2.0);
    ga->noise_std = (float)sqrt(sigma2 / 2.0);
    ga->Rr_root = (const float2*)gen->Rr_root;
    ga->Rt_root = (const float2*)gen->Rt_root;
    ga->real_roots = gen->real_roots != 0;
    if (!(gen->rho_t > -1.0 && gen->rho_t < 1.0 && gen->rho_r > -1.0 && gen->rho_r < 1.0)) { set_error("frame generation: |rho| must be below 1"); return AMPSM_EINVAL; }
    if ((gen->rho_t != 0.0 && gen->Rt_root) || (gen->rho_r != 0.0 && gen->Rr_root)) { set_error("frame generation: give rho or the root of a side, not both"); return AMPSM_EINVAL; }
    ga->ar_t = (float)gen->rho_t; ga->ar_t_c = (float)sqrt(1.0 - gen->rho_t * gen->rho_t);
    ga->ar_r = (float)gen->rho_r; ga->ar_r_c = (float)sqrt(1.0 - gen->rho_r * gen->rho_r);
    ga->K = a->K;
    for (int k = 0; k < AMPSM_MAX_K; ++k) {
        ga->sym[k] = make_float2(al.ref[k], al.imf[k]);
        ga->gray[k] = al.gray[k];
    }
    ga->x_out = (float2*)x;
    ga->idx_out = (long long*)idx;
    ga->sym_out = (long long*)sym;
    return 0;
}

int ampsm_generate_frames(const ampsm_problem* p, const ampsm_alphabet* a, const ampsm_gen* gen, int64_t frames, double sigma2, void* H,
                          void* y, void* x, int64_t* sym, int64_t* idx, void* stream) {
    Geom g{};
    GenArgs ga{};
    if (int e = make_gen(p, a, gen, sigma2, x, sym, idx, &g, &ga)) return e;
    if (frames < 0) { set_error("frame generation: frames < 0"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    return launch_generate_frames(ga, g, frames, (float2*)H, (float2*)y, (cudaStream_t)stream);
}

int ampsm_vamp_detect_generated(const ampsm_problem* p, const ampsm_alphabet* a, const ampsm_gen* gen, int64_t frames, double sigma2,
                                double sparsity, void* x, int64_t* sym, int64_t* idx, void* xmap, void* xmmse, float* var, int32_t* iters,
                                uint64_t* counters, void* workspace, void* stream) {
    Geom g{};
    GenArgs ga{};
    if (int e = make_gen(p, a, gen, sigma2, x, sym, idx, &g, &ga)) return e;
    if (frames < 0 || (frames > 0 && (!workspace || !x))) { set_error("VAMP on generated frames: workspace / ground-truth buffers are NULL or frames < 0"); return AMPSM_EINVAL; }
    if (p->R != p->n || p->n > p->N) { set_error("VAMP on generated frames: needs R = n <= N (got n=%d N=%d R=%d)", p->n, p->N, p->R); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    const size_t n = p->n, N = p->N;
    unsigned char* w = (unsigned char*)workspace;                      // same layout as ampsm_vamp_detect_from_h
    float2* yrot = (float2*)w;
    float* s = (float*)(w + align256((size_t)frames * n * 8));
    void* Vh = w + align256((size_t)frames * n * 8) + align256((size_t)frames * n * 4);
    float2* eye = (float2*)((unsigned char*)Vh + align256((size_t)frames * n * N * 8));
    if (int e = launch_svd_jacobi_gen(ga, g, frames, s, (float2*)Vh, yrot, (cudaStream_t)stream)) return e;
    if (int e = launch_identity(eye, p->n, (cudaStream_t)stream)) return e;
    return ampsm_vamp_detect(p, a, frames, 0, eye, 0, s, (int64_t)n, Vh, (int64_t)(n * N), yrot, sigma2, nullptr, sparsity, x, sym, idx,
                             xmap, xmmse, var, iters, nullptr, counters, stream);
}

// ---------------------------------------------------------------- SCAMP
int64_t ampsm_scamp_workspace_bytes(const ampsm_problem* p, int64_t frames) {
    Geom g{};
    DevAlphabet al{};
    ampsm_alphabet dummy{};
    dummy.K = 1;
    if (make_geom(p, &dummy, &g, &al, false)) return -1;
    return scamp_workspace_bytes(g, frames);
}

int ampsm_scamp_detect(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const float* W, const void* A,
                       const void* y, double sigma2, const float* sigma2_per_frame, const void* x_true,
                       const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse, float* psi, int32_t* iters,
                       float* traj, uint64_t* counters, void* workspace, void* stream) {
    ScampArgs k{};
    if (p && p->decision == 2) { set_error("SCAMP is defined for sectioned messages only (scamp.py:61-68)"); return AMPSM_EINVAL; }
    if (int e = make_geom(p, a, &k.g, &k.al, false)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || (frames > 0 && (!W || !A || !y))) { set_error("SCAMP: W / A / y is NULL or frames < 0"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    k.W = W; k.A = (const float2*)A; k.y = (const float2*)y; k.sigma2 = (float)sigma2; k.sigma2_pf = sigma2_per_frame;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = (float2*)xmap; k.xmmse = (float2*)xmmse; k.psi = psi; k.iters = iters; k.traj = traj; k.frames = frames;
    k.workspace = workspace;
    return launch_scamp(k, p->exp_f64 != 0, (cudaStream_t)stream);
}

int64_t ampsm_scamp_taps_workspace_bytes(const ampsm_problem* p, int64_t frames, int32_t Lh) {
    Geom g{};
    DevAlphabet al{};
    ampsm_alphabet dummy{};
    dummy.K = 1;
    if (make_geom(p, &dummy, &g, &al, false)) return -1;
    return scamp_taps_workspace_bytes(g, frames, Lh);
}

int ampsm_scamp_detect_taps(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const float* W, const void* taps, int32_t Lh,
                            const void* y, double sigma2, const float* sigma2_per_frame, const void* x_true,
                            const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse, float* psi, int32_t* iters,
                            float* traj, uint64_t* counters, void* workspace, void* stream) {
    ScampArgs k{};
    if (p && p->decision == 2) { set_error("SCAMP is defined for sectioned messages only (scamp.py:61-68)"); return AMPSM_EINVAL; }
    if (int e = make_geom(p, a, &k.g, &k.al, false)) return e;
    if (int e = check_loss_io(x_true, sym_true, idx_true)) return e;
    if (frames < 0 || Lh < 1 || (frames > 0 && (!W || !taps || !y))) { set_error("SCAMP taps: W / taps / y is NULL, Lh < 1 or frames < 0"); return AMPSM_EINVAL; }
    if (reinterpret_cast<uintptr_t>(taps) % 8) { set_error("SCAMP taps: taps must be 8-byte aligned"); return AMPSM_EINVAL; }
    if (frames == 0) return 0;
    k.W = W; k.A = nullptr; k.taps = (const float2*)taps; k.Lh = Lh; k.y = (const float2*)y; k.sigma2 = (float)sigma2; k.sigma2_pf = sigma2_per_frame;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.xmap = (float2*)xmap; k.xmmse = (float2*)xmmse; k.psi = psi; k.iters = iters; k.traj = traj; k.frames = frames;
    k.workspace = workspace;
    return launch_scamp(k, p->exp_f64 != 0, (cudaStream_t)stream);
}

int ampsm_scamp_detect_host(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const float* W, const void* A,
                            const void* y, double sigma2, const float* sigma2_per_frame, const void* x_true,
                            const int64_t* sym_true, const int64_t* idx_true, void* xmap, void* xmmse, float* psi,
                            int32_t* iters, float* traj, uint64_t* counters, int device) {
    if (!p || !a) { set_error("problem / alphabet pointer is NULL"); return AMPSM_EINVAL; }
    if (frames <= 0) return frames < 0 ? AMPSM_EINVAL : 0;
    const size_t L = (size_t)p->Na * p->Lin;
    std::vector<Field> f = {
        {y, nullptr, (size_t)p->n * 8},                        // 0
        {sigma2_per_frame, nullptr, 4},                        // 1
        {x_true, nullptr, (size_t)p->N * 8},                   // 2
        {sym_true, nullptr, L * 8},                            // 3
        {idx_true, nullptr, L * 8},                            // 4
        {nullptr, xmap, (size_t)p->N * 8},                     // 5
        {nullptr, xmmse, (size_t)p->N * 8},                    // 6
        {nullptr, psi, (size_t)p->Lin * 4},                    // 7
        {nullptr, iters, 4},                                   // 8
        {nullptr, traj, (size_t)p->max_iters * 12},            // 9
    };
    std::vector<std::pair<const void*, size_t>> shared = {{W, (size_t)p->Lout * p->Lin * 4}, {A, (size_t)p->n * p->N * 8}};
    std::vector<void*> sd;
    // workspace per frame: see ws_layout (Xh, Z, Zs, Xmap, scalars, denoiser scratch) -- bounded by this estimate
    const size_t ws_pf = (size_t)p->N * (8 + 8 + 24) + (size_t)p->n * 16 + (size_t)(p->Lin + p->Lout) * 8 + 64;
    const size_t ws_fixed = ((size_t)(p->n / 32 + 1) * (p->N / 32 + 1)) + 16 * 256;
    return run_host(device, frames, f, shared, &sd, ws_pf, ws_fixed, counters,
                    [&](long long f0, long long nf, void* const* d, unsigned long long* dc, void* scratch, cudaStream_t st) {
                        ampsm_problem q = *p;
                        q.frame_base = p->frame_base + f0;
                        if (ampsm_scamp_workspace_bytes(&q, nf) > (int64_t)(ws_fixed + (size_t)nf * ws_pf)) scratch = nullptr;
                        return ampsm_scamp_detect(&q, a, nf, (const float*)sd[0], sd[1], d[0], sigma2, (const float*)d[1], d[2],
                                                  (const int64_t*)d[3], (const int64_t*)d[4], d[5], d[6], (float*)d[7],
                                                  (int32_t*)d[8], (float*)d[9], (uint64_t*)dc, scratch, st);
                    });
}

// ---------------------------------------------------------------- Loss
int ampsm_loss_count(const ampsm_problem* p, const ampsm_alphabet* a, int64_t frames, const void* xmap, const void* xmmse,
                     const void* x_true, const int64_t* sym_true, const int64_t* idx_true, const int32_t* iters,
                     uint64_t* counters, void* stream) {
    LossArgs k{};
    if (int e = make_geom(p, a, &k.g, &k.al, false)) return e;
    if (!xmap || !xmmse || !x_true || !sym_true || !idx_true || !counters) { set_error("Loss: NULL pointer"); return AMPSM_EINVAL; }
    if (frames <= 0) return frames < 0 ? AMPSM_EINVAL : 0;
    k.xmap = (const float2*)xmap; k.xmmse = (const float2*)xmmse; k.iters = iters;
    k.io.x_true = (const float2*)x_true; k.io.sym_true = (const long long*)sym_true; k.io.idx_true = (const long long*)idx_true;
    k.io.counters = (unsigned long long*)counters;
    k.frames = frames;
    return launch_loss(k, (cudaStream_t)stream);
}

int ampsm_shrink(int kind, const ampsm_alphabet* a, double P0, double Ps, int64_t elems, int32_t M, const void* r,
                 const float* cov, int64_t cov_stride, void* out_c, float* out_f, double* der_sum, void* stream) {
    if (kind < AMPSM_SHRINK_BAYES || kind > AMPSM_SHRINK_SW_OOK) { set_error("Shrink: kind=%d outside 0..2", kind); return AMPSM_EINVAL; }
    if (!a || a->K < 1 || a->K > AMPSM_MAX_K) { set_error("Shrink: bad alphabet"); return AMPSM_EINVAL; }
    if (elems < 0 || (cov_stride != 0 && cov_stride != 1)) { set_error("Shrink: elems < 0 or cov_stride not in {0, 1}"); return AMPSM_EINVAL; }
    if (elems == 0) return 0;
    if (!r || !cov) { set_error("Shrink: r / cov is NULL"); return AMPSM_EINVAL; }
    if ((kind != AMPSM_SHRINK_OOK && !out_c) || (kind != AMPSM_SHRINK_BAYES && !out_f)) { set_error("Shrink: output pointer is NULL"); return AMPSM_EINVAL; }
    if (kind == AMPSM_SHRINK_SW_OOK && (M < 1 || elems % M != 0)) { set_error("Shrink: section size M=%d must divide elems", M); return AMPSM_EINVAL; }
    if (kind == AMPSM_SHRINK_OOK && !(Ps > 0.0)) { set_error("Shrink: Ps must be positive"); return AMPSM_EINVAL; }
    ShrinkArgs k{};
    k.al.K = a->K;
    for (int i = 0; i < a->K; ++i) { k.al.ref[i] = (float)a->re[i]; k.al.imf[i] = (float)a->im[i]; }
    k.P0 = (float)P0; k.Ps = (float)Ps;
    k.r = (const float2*)r; k.cov = cov; k.cov_stride = cov_stride; k.elems = elems; k.M = M;
    k.out_c = (float2*)out_c; k.out_f = out_f; k.sum = der_sum;
    return launch_shrink(k, kind, (cudaStream_t)stream);
}

}  // extern "C"
