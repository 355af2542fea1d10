"""ctypes binding of libampsm_b200.so (include/ampsm_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (plain ``nvcc -shared`` for sm_100a).  There is no CPU
fallback: ``lib()`` raises when the shared object is missing, and the detectors raise when no CUDA device is visible.
"""
import ctypes as C
import os

import numpy as np

MAX_K = 16
NUM_COUNTERS = 24
LIB_NAME = "libampsm_b200.so"
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", LIB_NAME)

COUNTER_NAMES = ['frames', 'frame_err', 'slot_err', 'slot_err_first', 'slot_err_mid', 'slot_err_last',
                 'index_err', 'symbol_err', 'index_bit_err', 'symbol_bit_err', 'iters', 'nan_frames']
SQERR_NAMES = ['sqerr', 'sqerr_first', 'sqerr_mid', 'sqerr_last']

# every symbol include/ampsm_b200.h declares (checked by tests/test_host_api.py::test_cabi_library_exports_every_declared_symbol)
EXPORTS = ["ampsm_version", "ampsm_last_error", "ampsm_device_info", "ampsm_bamp_detect", "ampsm_bamp_detect_taps", "ampsm_bamp_detect_host",
           "ampsm_vamp_detect", "ampsm_vamp_detect_host", "ampsm_svd_batched", "ampsm_vamp_from_h_workspace_bytes",
           "ampsm_vamp_detect_from_h", "ampsm_scamp_workspace_bytes", "ampsm_scamp_detect", "ampsm_scamp_taps_workspace_bytes", "ampsm_scamp_detect_taps",
           "ampsm_scamp_detect_host", "ampsm_loss_count", "ampsm_shrink", "ampsm_probe_fp32_tflops", "ampsm_probe_fp32x2_tflops", "ampsm_probe_fp64_tflops",
           "ampsm_launch_count", "ampsm_host_alloc", "ampsm_host_free", "ampsm_host_numa_info", "ampsm_generate_frames",
           "ampsm_vamp_detect_generated", "ampsm_vamp2_detect"]


class Alphabet(C.Structure):
    _fields_ = [("K", C.c_int32), ("gray", C.c_int32 * MAX_K), ("re", C.c_double * MAX_K), ("im", C.c_double * MAX_K)]


class Problem(C.Structure):
    _fields_ = [("n", C.c_int32), ("N", C.c_int32), ("R", C.c_int32), ("Nt", C.c_int32), ("Na", C.c_int32),
                ("Nr", C.c_int32), ("Lin", C.c_int32), ("Lout", C.c_int32), ("max_iters", C.c_int32),
                ("early_exit", C.c_int32), ("shift_mode", C.c_int32), ("exp_f64", C.c_int32), ("decision", C.c_int32),
                ("index_bits_kept", C.c_int32), ("kernel", C.c_int32), ("reserved0", C.c_int32), ("frame_base", C.c_int64)]


class Gen(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("counter_base", C.c_int64), ("h_var", C.c_double), ("Rr_root", C.c_void_p), ("Rt_root", C.c_void_p),
                ("real_roots", C.c_int32), ("reserved0", C.c_int32), ("rho_r", C.c_double), ("rho_t", C.c_double)]


class AmpsmError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the CUDA library; fail loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("AMPSM_LIB", LIB_PATH)        # development builds of the same library (scripts/build_clk.sh)
    if not os.path.exists(path):
        raise AmpsmError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                         "(nvcc, sm_100a).  There is no CPU fallback for the detector hot path.")
    L = C.CDLL(path)
    vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
    PP, AP = C.POINTER(Problem), C.POINTER(Alphabet)
    L.ampsm_version.restype = C.c_char_p
    L.ampsm_last_error.restype = C.c_char_p
    L.ampsm_device_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    bamp = [PP, AP, i64, vp, i64, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_bamp_detect.argtypes = bamp + [vp]
    L.ampsm_bamp_detect_host.argtypes = bamp + [i32]
    L.ampsm_bamp_detect_taps.argtypes = [PP, AP, i64, vp, i64, i32, i32] + bamp[5:] + [vp]
    vamp = [PP, AP, i64, i32, vp, i64, vp, i64, vp, i64, vp, dbl, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_vamp_detect.argtypes = vamp + [vp]
    L.ampsm_vamp_detect_host.argtypes = vamp + [i32]
    L.ampsm_vamp2_detect.argtypes = [PP, AP, i64, vp, i64, vp, i64, vp, i64, vp, dbl, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_svd_batched.argtypes = [i64, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ampsm_vamp_from_h_workspace_bytes.argtypes = [PP, i64]
    L.ampsm_vamp_from_h_workspace_bytes.restype = i64
    L.ampsm_vamp_detect_from_h.argtypes = [PP, AP, i64, vp, vp, dbl, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    GP = C.POINTER(Gen)
    L.ampsm_generate_frames.argtypes = [PP, AP, GP, i64, dbl, vp, vp, vp, vp, vp, vp]
    L.ampsm_vamp_detect_generated.argtypes = [PP, AP, GP, i64, dbl, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    scamp = [PP, AP, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_scamp_detect.argtypes = scamp + [vp, vp]
    L.ampsm_scamp_detect_host.argtypes = scamp + [i32]
    L.ampsm_scamp_workspace_bytes.argtypes = [PP, i64]
    L.ampsm_scamp_workspace_bytes.restype = i64
    L.ampsm_scamp_taps_workspace_bytes.argtypes = [PP, i64, i32]
    L.ampsm_scamp_taps_workspace_bytes.restype = i64
    L.ampsm_scamp_detect_taps.argtypes = [PP, AP, i64, vp, vp, i32, vp, dbl, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_loss_count.argtypes = [PP, AP, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ampsm_shrink.argtypes = [i32, AP, dbl, dbl, i64, i32, vp, vp, i64, vp, vp, vp, vp]
    L.ampsm_probe_fp32_tflops.argtypes = [i32, C.POINTER(dbl)]
    L.ampsm_probe_fp32x2_tflops.argtypes = [i32, C.POINTER(dbl)]
    L.ampsm_probe_fp64_tflops.argtypes = [i32, C.POINTER(dbl)]
    L.ampsm_host_alloc.argtypes = [i32, C.c_size_t, C.POINTER(vp)]
    L.ampsm_host_free.argtypes = [vp]
    L.ampsm_host_numa_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32)]
    L.ampsm_launch_count.argtypes = [i32]
    L.ampsm_launch_count.restype = i64
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        msg = lib().ampsm_last_error().decode(errors="replace")
        raise AmpsmError(f"{what} failed (code {rc}): {msg}")


def make_alphabet(config):
    a = Alphabet()
    sym = np.asarray(config.symbols, dtype=np.complex128)
    if sym.shape[0] > MAX_K:
        raise AmpsmError(f"alphabet of {sym.shape[0]} symbols exceeds AMPSM_MAX_K={MAX_K}")
    a.K = sym.shape[0]
    for k in range(a.K):
        a.gray[k] = int(config.gray[k])
        a.re[k] = float(sym[k].real)
        a.im[k] = float(sym[k].imag)
    return a


def make_problem(config, frames, *, R=0, early_exit=True, shift='section', exp='f32', kernel='auto', frame_base=0,
                 max_iters=None):
    """Geometry + switches for one call of `frames` frames (index-bit truncation as loss.py:20 with B=frames)."""
    import math
    p = Problem()
    p.n, p.N, p.R = config.Nr * config.Lout, config.Nt * config.Lin, R
    p.Nt, p.Na, p.Nr, p.Lin, p.Lout = config.Nt, config.Na, config.Nr, config.Lin, config.Lout
    p.max_iters = int(max_iters if max_iters is not None else config.N_Layers)
    p.early_exit = 1 if early_exit else 0
    p.shift_mode = {'section': 0, 'reference': 1}[shift]
    p.exp_f64 = {'f32': 0, 'f64': 1}[exp]
    if p.shift_mode == 1:
        p.exp_f64 = 1
    p.decision = {'sparc': 0, 'segmented': 1, 'random': 2}[config.mode]      # decision rule and, for 'random', the i.i.d. prior
    count = config.Lin * max(int(frames), 1) * config.Na
    p.index_bits_kept = int(math.ceil(math.log2(count))) if count > 0 else 0
    p.kernel = {'auto': 0, 'generic': 1, 'fast': 2}[kernel]
    p.frame_base = int(frame_base)
    return p


def counters_to_dict(buf):
    """buf: 24 x int64 (numpy) as written by the library."""
    buf = np.ascontiguousarray(buf, dtype=np.int64)
    out = {k: int(buf[i]) for i, k in enumerate(COUNTER_NAMES)}
    sq = buf[16:20].view(np.float64)
    out.update({k: float(sq[i]) for i, k in enumerate(SQERR_NAMES)})
    return out


class HostBuffer:
    """Pinned host memory placed on the NUMA node of a GPU (``ampsm_host_alloc``), viewed as a numpy array; for the buffers
    handed to the ``*_detect_host`` entry points."""

    def __init__(self, device: int, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(v) for v in np.atleast_1d(shape))
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        check(lib().ampsm_host_alloc(int(device), self.nbytes, C.byref(p)), "ampsm_host_alloc")
        self.ptr = p.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(self.shape))).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            lib().ampsm_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
