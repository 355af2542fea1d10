"""On-device generation of Monte-Carlo frames (csrc/framegen.cuh): the throughput-sweep replacement of the reference's
per-epoch draws -- ``Channel.generate_channel`` (channel.py:53-55), ``Data.generate_message`` in its sectioned modes
(data.py:74-91) and ``Channel.awgn`` (channel.py:113-115) -- with a counter-based Philox4x32-10 stream, one kernel, one warp
per frame.  ``FrameStream.frames`` writes H, y, x and the labels to HBM (for BAMP and for checking); ``VAMP.detect_generated``
draws the same frames inside the Jacobi SVD kernel, so the channel matrix never exists in HBM.
The draws are not the reference's numpy / torch sequences: parity subsets keep the reference's own RNG path."""
import numpy as np
import torch

from . import _cabi
from .config import Config


def exp_corr_root(size: int, rho: float, device) -> torch.Tensor:
    """Hermitian square root of the exponential correlation matrix R[i, j] = rho^|i-j| (BASELINE config 5) as complex64."""
    i = np.arange(size)
    R = float(rho) ** np.abs(i[:, None] - i[None, :])
    w, V = np.linalg.eigh(R)
    root = (V * np.sqrt(np.clip(w, 0.0, None))) @ V.T
    return torch.as_tensor(root.astype(np.complex64), device=device).contiguous()


class FrameStream:
    """A reproducible stream of frames for ``config`` (Lin = Lh = 1, sectioned messages): frame ``k`` of the stream is the same
    whatever chunk, shard or call draws it."""

    def __init__(self, config: Config, seed: int = 0, channel='iid', rho_t=0.0, rho_r=0.0, device=None, method='ar1'):
        if config.Lin != 1 or config.Lh != 1 or config.mode == 'random':
            raise _cabi.AmpsmError("FrameStream draws sectioned messages on memoryless channels (Lin = Lh = 1)")
        self.config = config
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.device = torch.device(device if device is not None else config.device)
        if self.device.type != "cuda":
            raise _cabi.AmpsmError("FrameStream runs on the GPU only (no CPU fallback)")
        self.Rr = self.Rt = None
        self.real_roots = False
        self.rho_r = self.rho_t = 0.0
        if channel == 'kronecker' and method == 'ar1':
            # exponential correlation by two AR(1) recursions inside the kernel: H = A G B with A A^H = Rr, B^H B = Rt, the
            # distribution of Rr^(1/2) G Rt^(1/2) at a sixtieth of the arithmetic
            self.rho_r, self.rho_t = float(rho_r), float(rho_t)
        elif channel == 'kronecker' and method == 'roots':
            self.Rr = exp_corr_root(config.n, rho_r, self.device)
            self.Rt = exp_corr_root(config.N, rho_t, self.device)
            self.real_roots = True                              # rho is real: so are the roots
        elif channel == 'kronecker':
            raise ValueError("method is 'ar1' or 'roots'")
        elif channel not in ('iid', 'kronecker'):
            raise ValueError(f"unknown channel model {channel!r}")
        self._alphabet = _cabi.make_alphabet(config)

    def gen_struct(self, first_frame: int) -> _cabi.Gen:
        g = _cabi.Gen()
        g.seed, g.counter_base, g.h_var = self.seed, int(first_frame), 1.0 / self.config.Nr
        g.Rr_root = self.Rr.data_ptr() if self.Rr is not None else None
        g.Rt_root = self.Rt.data_ptr() if self.Rt is not None else None
        g.real_roots = 1 if self.real_roots else 0
        g.rho_r, g.rho_t = self.rho_r, self.rho_t
        return g

    def truth_buffers(self, frames: int):
        cfg, dev = self.config, self.device
        return (torch.empty(frames, cfg.N, dtype=torch.complex64, device=dev), torch.empty(frames * cfg.L, dtype=torch.int64, device=dev),
                torch.empty(frames * cfg.L, dtype=torch.int64, device=dev))

    def frames(self, first_frame: int, frames: int, snr: float, frame_base: int = 0, with_channel=True):
        """Frames ``first_frame .. first_frame + frames - 1`` of the stream at linear SNR ``snr`` (sigma^2 = (Na / Nr) / snr,
        bamp.py:111,134): ``(H (F, n, N), y (F, n), x (F, N), Gray labels (F L,), flat positions (F L,))`` like
        ``simulate.device_frames``; positions count from ``frame_base``."""
        cfg, dev = self.config, self.device
        H = torch.empty(frames, cfg.n, cfg.N, dtype=torch.complex64, device=dev) if with_channel else None
        y = torch.empty(frames, cfg.n, dtype=torch.complex64, device=dev)
        x, sym, idx = self.truth_buffers(frames)
        prob = _cabi.make_problem(cfg, frames, frame_base=frame_base)
        gen = self.gen_struct(first_frame)
        with torch.cuda.device(dev):
            rc = _cabi.lib().ampsm_generate_frames(prob, self._alphabet, gen, frames, float((cfg.Na / cfg.Nr) / snr),
                                                   H.data_ptr() if H is not None else None, y.data_ptr(), x.data_ptr(), sym.data_ptr(),
                                                   idx.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
        _cabi.check(rc, "ampsm_generate_frames")
        return H, y, x, sym, idx
