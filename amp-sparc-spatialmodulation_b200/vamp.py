"""VAMP detector with the reference call signature (/root/reference/vamp.py:151-191), running on sm_100a.

``VAMP(config)(U, s, Vh, y, SNR, x, symbols, indices) -> Loss`` with the caller's thin SVD (vamp_model.py:58).
Factors may be shared by the call -- U (n, R), s (R,), Vh (R, N) -- or given per frame with a leading frame
axis.  complex128 factors select the float64 linear stage (the reference fed with upcast inputs); ``xmmse`` /
``var`` are complex64 / float32 in both cases (vamp.py:119).  Loss is fed ``r`` as xmap (vamp.py:187).
"""
import torch

from . import _cabi
from ._detect import Detection, Detector, ptr, dense


class VAMP(Detector):
    def __init__(self, config, **kw) -> None:
        super().__init__(config, **kw)
        self.sparsity = config.Na / config.Nt              # vamp.py:155

    def detect(self, U, s, Vh, y, SNR, x=None, symbols=None, indices=None, frame_base=0) -> Detection:
        dev = self._cuda_device(y, Vh, U)
        cfg = self.config
        n, N = cfg.Nr * cfg.Lout, cfg.Nt * cfg.Lin
        dbl = Vh.dtype == torch.complex128
        ct, rt = (torch.complex128, torch.float64) if dbl else (torch.complex64, torch.float32)
        y = dense(y, dev, ct, -1, n)
        F = y.shape[0]
        U, s, Vh = dense(U, dev, ct), dense(s, dev, rt), dense(Vh, dev, ct)
        R = Vh.shape[-2]
        if tuple(Vh.shape[-2:]) != (R, N) or tuple(U.shape[-2:]) != (n, R) or s.shape[-1] != R:
            raise RuntimeError(f"factor shapes U{tuple(U.shape)} s{tuple(s.shape)} Vh{tuple(Vh.shape)} do not match n={n}, N={N}")
        def stride(t, base_dim, size):
            if t.dim() == base_dim:
                return 0
            if t.dim() == base_dim + 1 and t.shape[0] == F:
                return size
            raise RuntimeError(f"factor with shape {tuple(t.shape)} is neither shared nor per-frame for {F} frames")
        sU, ss, sV = stride(U, 2, n * R), stride(s, 1, R), stride(Vh, 2, R * N)
        xt = None if x is None else dense(x, dev, torch.complex64, -1, N)
        sym, idx = self._labels(symbols, indices, dev) if xt is not None else (None, None)
        counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
        iters = torch.empty(F, dtype=torch.int32, device=dev)
        xmap = torch.empty(F, N, 1, dtype=ct, device=dev) if self.outputs else None
        xmmse = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        var = torch.empty(F, N, 1, dtype=torch.float32, device=dev) if self.outputs else None
        traj = torch.empty(F, cfg.N_Layers, 3, dtype=torch.float32, device=dev) if self.trajectory else None
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_vamp_detect(
                self._problem(F, R=R, frame_base=frame_base), self._alphabet, F, 1 if dbl else 0, U.data_ptr(), sU,
                s.data_ptr(), ss, Vh.data_ptr(), sV, y.data_ptr(), float(self.E / SNR), None, float(self.sparsity),
                ptr(xt), ptr(sym), ptr(idx), ptr(xmap), ptr(xmmse), ptr(var), iters.data_ptr(), ptr(traj),
                counters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_vamp_detect")
        return Detection(F, counters, iters, xmap, xmmse, var, traj)

    def forward(self, U, s, Vh, y, SNR, x, symbols, indices):
        return self._wrap(self.detect(U, s, Vh, y, SNR, x, symbols, indices))

    def detect_from_channel(self, H, y, SNR, x=None, symbols=None, indices=None, frame_base=0) -> Detection:
        """vamp_model.py:56-61 for a batch of frames with their own channel matrices: batched Jacobi SVD on the device
        (csrc/svd_jacobi.cu) followed by the VAMP iterations, one C-ABI call (``ampsm_vamp_detect_from_h``)."""
        dev = self._cuda_device(y, H)
        cfg = self.config
        n, N = cfg.Nr * cfg.Lout, cfg.Nt * cfg.Lin
        y = dense(y, dev, torch.complex64, -1, n)
        F = y.shape[0]
        H = dense(H, dev, torch.complex64)
        if tuple(H.shape) != (F, n, N):
            raise RuntimeError(f"H{tuple(H.shape)} is not ({F}, {n}, {N})")
        xt = None if x is None else dense(x, dev, torch.complex64, -1, N)
        sym, idx = self._labels(symbols, indices, dev) if xt is not None else (None, None)
        counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
        iters = torch.empty(F, dtype=torch.int32, device=dev)
        xmap = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        xmmse = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        var = torch.empty(F, N, 1, dtype=torch.float32, device=dev) if self.outputs else None
        traj = torch.empty(F, cfg.N_Layers, 3, dtype=torch.float32, device=dev) if self.trajectory else None
        prob = self._problem(F, R=min(n, N), frame_base=frame_base)
        ws = torch.empty(int(_cabi.lib().ampsm_vamp_from_h_workspace_bytes(prob, F)), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_vamp_detect_from_h(
                prob, self._alphabet, F, H.data_ptr(), y.data_ptr(), float(self.E / SNR), None, float(self.sparsity),
                ptr(xt), ptr(sym), ptr(idx), ptr(xmap), ptr(xmmse), ptr(var), iters.data_ptr(), ptr(traj),
                counters.data_ptr(), ws.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_vamp_detect_from_h")
        return Detection(F, counters, iters, xmap, xmmse, var, traj)

    def detect_generated(self, stream, first_frame: int, frames: int, SNR, frame_base=0, return_truth=False):
        """vamp_model.py:44-61 for ``frames`` frames of a ``framegen.FrameStream``: the Jacobi SVD kernel draws every frame
        (channel, message, noise) straight into its shared-memory tile, factorises it and feeds the VAMP iterations -- the
        channel matrix never exists in HBM (``ampsm_vamp_detect_generated``).  Same frames, bit for bit, as
        ``stream.frames(first_frame, frames, SNR)`` followed by ``detect_from_channel``."""
        cfg, dev = self.config, stream.device
        if not torch.cuda.is_available():
            raise _cabi.AmpsmError("no CUDA device: the detector hot path has no CPU fallback")
        n, N, F = cfg.Nr * cfg.Lout, cfg.Nt * cfg.Lin, int(frames)
        x, sym, idx = stream.truth_buffers(F)
        counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
        iters = torch.empty(F, dtype=torch.int32, device=dev)
        xmap = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        xmmse = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        var = torch.empty(F, N, 1, dtype=torch.float32, device=dev) if self.outputs else None
        prob = self._problem(F, R=min(n, N), frame_base=frame_base)
        ws = torch.empty(int(_cabi.lib().ampsm_vamp_from_h_workspace_bytes(prob, F)), dtype=torch.uint8, device=dev)
        gen = stream.gen_struct(first_frame)
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_vamp_detect_generated(
                prob, self._alphabet, gen, F, float(self.E / SNR), float(self.sparsity), x.data_ptr(), sym.data_ptr(), idx.data_ptr(),
                ptr(xmap), ptr(xmmse), ptr(var), iters.data_ptr(), counters.data_ptr(), ws.data_ptr(),
                torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_vamp_detect_generated")
        det = Detection(F, counters, iters, xmap, xmmse, var, None)
        return (det, x, sym, idx) if return_truth else det


def svd_batched(H, return_sweeps=False):
    """Thin SVD of a batch of wide complex64 matrices on the device, ``H (F, n, N) -> U (F, n, n), s (F, n), Vh (F, n, N)``
    with ``s`` descending -- the device-side stand-in for ``torch.linalg.svd(A, full_matrices=False)`` at
    vamp_model.py:58 (one-sided Jacobi, one warp per matrix; singular-vector phases differ from LAPACK's)."""
    if not H.is_cuda:
        raise _cabi.AmpsmError("svd_batched runs on the GPU only (no CPU fallback)")
    H = H.to(torch.complex64).resolve_conj().contiguous()
    F, n, N = H.shape
    U = torch.empty(F, n, n, dtype=torch.complex64, device=H.device)
    s = torch.empty(F, n, dtype=torch.float32, device=H.device)
    Vh = torch.empty(F, n, N, dtype=torch.complex64, device=H.device)
    sw = torch.empty(F, dtype=torch.int32, device=H.device) if return_sweeps else None
    with torch.cuda.device(H.device):
        rc = _cabi.lib().ampsm_svd_batched(F, n, N, H.data_ptr(), U.data_ptr(), s.data_ptr(), Vh.data_ptr(), ptr(sw),
                                           torch.cuda.current_stream(H.device).cuda_stream)
    _cabi.check(rc, "ampsm_svd_batched")
    return (U, s, Vh, sw) if return_sweeps else (U, s, Vh)
