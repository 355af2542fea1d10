"""BAMP detector with the reference call signature (/root/reference/bamp.py:104-143), running on sm_100a.

``BAMP(config)(H, y, SNR, x, symbols, indices) -> Loss``.  One call may hold many frames: ``y`` is (F, n, 1) and
``H`` is either one (n, N) matrix shared by the call (the reference's layout) or (F, n, N), one matrix per frame.
Every frame is processed exactly as a ``batch=1`` reference call (per-frame soft-max shift and exit test,
SURVEY.md App. B.1).  All iterations, the denoiser, the hard decision and the error counters run inside one
persistent CUDA kernel (csrc/bamp_generic.cu, csrc/bamp_fast.cu) reached through ``ampsm_bamp_detect``.
"""
import torch

from . import _cabi
from ._detect import Detection, Detector, ptr


class BAMP(Detector):
    def detect(self, H, y, SNR, x=None, symbols=None, indices=None, frame_base=0) -> Detection:
        """Enqueue the kernel on the current stream and return device-side results without synchronising."""
        dev = self._cuda_device(y, H)
        cfg = self.config
        n, N = cfg.Nr * cfg.Lout, cfg.Nt * cfg.Lin
        y = y.to(dev, torch.complex64).reshape(-1, n).contiguous()
        F = y.shape[0]
        H = H.to(dev, torch.complex64).contiguous()
        if H.dim() == 2:
            stride = 0
        elif H.dim() == 3 and H.shape[0] == F:
            stride = n * N
        else:
            raise RuntimeError(f"H must be (n, N) or (frames, n, N); got {tuple(H.shape)} for {F} frames")
        if tuple(H.shape[-2:]) != (n, N):
            raise RuntimeError(f"H has shape {tuple(H.shape)}, expected (..., {n}, {N})")
        xt = None if x is None else x.to(dev, torch.complex64).reshape(-1, N).contiguous()
        sym, idx = self._labels(symbols, indices, dev) if xt is not None else (None, None)
        counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
        iters = torch.empty(F, dtype=torch.int32, device=dev)
        xmap = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        xmmse = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        var = torch.empty(F, N, 1, dtype=torch.float32, device=dev) if self.outputs else None
        traj = torch.empty(F, cfg.N_Layers, 3, dtype=torch.float32, device=dev) if self.trajectory else None
        sigma2 = self.E / SNR                                # bamp.py:134
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_bamp_detect(
                self._problem(F, frame_base=frame_base), self._alphabet, F, H.data_ptr(), stride, y.data_ptr(),
                float(sigma2), None, ptr(xt), ptr(sym), ptr(idx), ptr(xmap), ptr(xmmse), ptr(var), iters.data_ptr(),
                ptr(traj), counters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_bamp_detect")
        return Detection(F, counters, iters, xmap, xmmse, var, traj)

    def forward(self, H, y, SNR, x, symbols, indices):
        return self._wrap(self.detect(H, y, SNR, x, symbols, indices))
