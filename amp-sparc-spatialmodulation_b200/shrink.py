"""``Shrink`` -- element-wise denoiser family of the reference (/root/reference/shrink.py:8-166) on sm_100a.

``Shrink(config, fn)(r, cov)`` with the reference's return conventions: ``'bayes'`` -> posterior mean (complex64);
``'shrinkOOK'`` -> ``(exp, dxdr)`` with ``dxdr`` the 0-dim batch mean of the derivative; ``sw_shrinkOOK(r, cov)`` ->
``(Exp complex64, Var float32)`` per section.  ``'shrink'`` and ``'lasso'`` are accepted by the constructor like in the
reference and fail when called like in the reference (``UnboundLocalError`` at shrink.py:113, ``AttributeError`` at
shrink.py:135).  Kernels: csrc/shrink.cu through ``ampsm_shrink``; no CPU fallback.
"""
import torch
from torch import nn

from . import _cabi
from ._detect import Detector, dense
from .config import Config

_KINDS = ("bayes", "shrink", "lasso", "shrinkOOK")


class Shrink(nn.Module):
    def __init__(self, config: Config, shrink_fn: str) -> None:
        super().__init__()
        assert shrink_fn in _KINDS, "shrink_fn needs to be one of " + ", ".join(_KINDS)
        self.config, self.kind = config, shrink_fn
        self.Ps, self.P0 = float(config.Ps), float(config.P0)
        self.M = config.Nt // config.Na                      # shrink.py:30-32
        self.L = config.Na * config.Lin
        self.B = config.B
        self._alphabet = _cabi.make_alphabet(config)

    def _run(self, kind, r, cov):
        dev = Detector._cuda_device(r, cov)
        shape = tuple(r.shape)
        rc64 = dense(r, dev, torch.complex64)
        cov = torch.as_tensor(cov, dtype=torch.float32).to(dev)
        if cov.numel() == 1:
            stride, covf = 0, cov.reshape(1).contiguous()
        else:
            stride, covf = 1, cov.expand(shape).contiguous()
        elems = rc64.numel()
        out_c = torch.empty(shape, dtype=torch.complex64, device=dev) if kind != 1 else None
        out_f = torch.empty(shape, dtype=torch.float32, device=dev) if kind != 0 else None
        acc = torch.zeros(1, dtype=torch.float64, device=dev) if kind == 1 else None
        p = lambda t: None if t is None else t.data_ptr()
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_shrink(kind, self._alphabet, self.P0, self.Ps, elems, self.M, rc64.data_ptr(),
                                          covf.data_ptr(), stride, p(out_c), p(out_f), p(acc),
                                          torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_shrink")
        return out_c, out_f, acc, elems

    def forward(self, r, cov):
        if self.kind == "bayes":
            return self._run(0, r, cov)[0]
        if self.kind == "shrinkOOK":
            _, e, acc, elems = self._run(1, r, cov)
            return e, (acc[0] / elems).to(torch.float32)
        if self.kind == "shrink":
            raise UnboundLocalError("cannot access local variable 'd0' where it is not associated with a value "
                                    "(the reference's Shrink.shrink fails the same way, shrink.py:113)")
        raise AttributeError("'Shrink' object has no attribute 'lmda' (as the reference, shrink.py:135)")

    def sw_shrinkOOK(self, r, cov):
        out_c, out_f, _, _ = self._run(2, r, cov)
        return out_c.view(self.B, self.L * self.M, 1), out_f.view(self.B, self.L * self.M, 1)
