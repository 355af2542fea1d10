"""The reference's second VAMP with its call signature (/root/reference/vamp2.py:97-131), running on sm_100a.

``vamp2.VAMP(config, damping)(U, s, Vh, y, SNR, x, symbols, indices) -> Loss`` -- "direct implementation of Rangan (with
damping)": a precision-like scalar ``gamma`` in place of vamp.py's variance pair, the posterior mean damped by ``rho`` between
iterations, the variance of the denoiser as ``E|s|^2 - |E s|^2``.  None of the reference's drivers imports the module; it is
kept as a drop-in for callers that do (``from vamp2 import VAMP``).  complex64 factors, shared by the call or per frame, as
``vamp.VAMP``; ``trajectory=True`` returns ``{gamma, mean var, mse}`` per iteration.  One CTA per frame (csrc/vamp2.cu).
"""
import torch

from . import _cabi
from ._detect import Detection, Detector, ptr, dense


class VAMP(Detector):
    def __init__(self, config, damping: float = 1.0, **kw) -> None:
        super().__init__(config, **kw)
        if config.mode == 'random':
            raise _cabi.AmpsmError("vamp2 with generator_mode='random' denoises with Shrink('bayes') (vamp2.py:45-46): use Shrink; "
                                   "the detector kernel runs the sectioned modes")
        self.damping = float(damping)                      # vamp2.py:98-102: every layer gets the same rho

    def detect(self, U, s, Vh, y, SNR, x=None, symbols=None, indices=None, frame_base=0) -> Detection:
        dev = self._cuda_device(y, Vh, U)
        cfg = self.config
        n, N = cfg.Nr * cfg.Lout, cfg.Nt * cfg.Lin
        y = dense(y, dev, torch.complex64, -1, n)
        F = y.shape[0]
        U, s, Vh = dense(U, dev, torch.complex64), dense(s, dev, torch.float32), dense(Vh, dev, torch.complex64)
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
        xmap = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        xmmse = torch.empty(F, N, 1, dtype=torch.complex64, device=dev) if self.outputs else None
        var = torch.empty(F, N, 1, dtype=torch.float32, device=dev) if self.outputs else None
        traj = torch.empty(F, cfg.N_Layers, 3, dtype=torch.float32, device=dev) if self.trajectory else None
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_vamp2_detect(
                self._problem(F, R=R, frame_base=frame_base), self._alphabet, F, U.data_ptr(), sU, s.data_ptr(), ss, Vh.data_ptr(), sV,
                y.data_ptr(), float(self.E / SNR), None, self.damping, ptr(xt), ptr(sym), ptr(idx), ptr(xmap), ptr(xmmse), ptr(var),
                iters.data_ptr(), ptr(traj), counters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_vamp2_detect")
        return Detection(F, counters, iters, xmap, xmmse, var, traj)

    def forward(self, U, s, Vh, y, SNR, x, symbols, indices):
        return self._wrap(self.detect(U, s, Vh, y, SNR, x, symbols, indices))
