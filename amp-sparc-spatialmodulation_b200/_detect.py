"""Plumbing shared by the BAMP / SCAMP / VAMP modules: device placement, label upload, result wrapping."""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch
from torch import nn

from . import _cabi
from .config import Config
from .loss import Loss
from ._tensors import dense, ptr  # noqa: F401  (re-exported for the detector modules)


@dataclass
class Detection:
    """Device-side result of one detector call (an addition to the reference surface; nothing here syncs)."""
    frames: int
    counters: torch.Tensor                 # int64[24] on the device, layout of include/ampsm_b200.h
    iters: torch.Tensor                    # int32[frames]
    xmap: Optional[torch.Tensor] = None    # (frames, N, 1)
    xmmse: Optional[torch.Tensor] = None   # (frames, N, 1) complex64
    var: Optional[torch.Tensor] = None     # BAMP / VAMP posterior variance, or SCAMP psi
    traj: Optional[torch.Tensor] = None    # (frames, max_iters, 3) float32

    def counters_dict(self):
        return _cabi.counters_to_dict(self.counters.cpu().numpy())


class Detector(nn.Module):
    """Common constructor switches.

    early_exit : per-frame ``torch.allclose`` exit as the reference (bamp.py:140) or exactly N_Layers iterations
    shift      : 'section' (per-section soft-max shift, finite everywhere) or 'reference' (frame-global max|x| in
                 float64 as bamp.py:70 -- reproduces the reference's NaN frames)
    exp        : 'f32' or 'f64' arithmetic for the denoiser exponents (the reference uses float64)
    kernel     : 'auto' | 'generic' | 'fast' (the register-resident kernels; raises when the shape has none)
    trajectory : also return per-iteration {tau, var, mse} means (costs a few block reductions per iteration)
    outputs    : materialise xmap / xmmse in HBM (needed by ``last``); the counters never need them
    """

    def __init__(self, config: Config, early_exit=True, shift='section', exp='f32', kernel='auto', trajectory=False,
                 outputs=True) -> None:
        super().__init__()
        self.config = config
        self.E = config.Na / config.Nr                      # bamp.py:111
        self.L = Loss(config)
        self.early_exit, self.shift, self.exp, self.kernel = early_exit, shift, exp, kernel
        self.trajectory, self.outputs = trajectory, outputs
        self.last: Optional[Detection] = None
        self._alphabet = _cabi.make_alphabet(config)

    # -- helpers -----------------------------------------------------------------------------------------------
    @staticmethod
    def _cuda_device(*tensors):
        if not torch.cuda.is_available():
            raise _cabi.AmpsmError("no CUDA device: the detectors run sm_100a kernels only (no CPU fallback)")
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                return t.device
        return torch.device('cuda', torch.cuda.current_device())

    @staticmethod
    def _labels(symbols, indices, dev):
        if symbols is None or indices is None:
            return None, None
        def up(v):
            if isinstance(v, torch.Tensor):
                return v.to(dev, torch.int64).contiguous()
            return torch.as_tensor(np.ascontiguousarray(v, dtype=np.int64)).to(dev)
        return up(symbols), up(indices)

    def _problem(self, frames, **kw):
        opts = dict(early_exit=self.early_exit, shift=self.shift, exp=self.exp, kernel=self.kernel)
        opts.update(kw)
        return _cabi.make_problem(self.config, frames, **opts)

    def _wrap(self, det: Detection) -> Loss:
        """Reference behaviour of forward(): reset the module's own Loss and fill it (bamp.py:135,142)."""
        self.last = det
        c = det.counters_dict()                             # the one host sync of the call
        self.L.dump()
        # 'T' is the exit iteration of the call; with per-frame exits it is the mean over the frames of the call
        T = c['iters'] / max(c['frames'], 1)
        self.L.record(c, int(T) if float(T).is_integer() else T)
        return self.L
