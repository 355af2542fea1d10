"""Dry-run of the Python host paths on a machine WITHOUT a GPU: every C-ABI call is replaced by a stub that only checks the
argument count against the ctypes prototype.  Catches naming / arity mistakes before GPU time is spent; computes nothing."""
import contextlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200 import _cabi, _detect  # noqa: E402

REAL = _cabi.lib()


class Stub:
    def __getattr__(self, name):
        real = getattr(REAL, name)

        def call(*a):
            assert len(a) == len(real.argtypes), (name, len(a), len(real.argtypes))
            return 0
        return call


class _Stream:
    cuda_stream = 0


_cabi.lib = lambda: Stub()
_detect.Detector._cuda_device = staticmethod(lambda *t: torch.device('cpu'))
torch.cuda.current_stream = lambda d=None: _Stream()
torch.cuda.device = lambda d: contextlib.nullcontext()

if __name__ == "__main__":
    cfg = pkg.Config(6, 2, 4, 5, 3, batch=2, generator_mode='sparc', alphabet='QPSK', channel_profile='uniform',
                     channel_truncation='cyclic', device='cpu')
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    H, taps = ch.generate_channel(return_taps=True)
    x, s, i = da.generate_message()
    y = H @ x
    amp = pkg.BAMP(cfg, trajectory=True)
    amp.detect_taps(taps, y, 3.0, x, s, i, cyclic=True)
    amp.detect(H, y, 3.0, x, s, i)
    pkg.BAMP(cfg, structured=False).detect(H, y, 3.0, x, s, i)
    pkg.Shrink(cfg, 'bayes')(x, torch.tensor(0.3))
    W, A = ch.generate_as_sparc()
    U, sv, Vh = torch.linalg.svd(A, full_matrices=False)
    pkg.VAMP(cfg).detect(U, sv, Vh, A @ x, 3.0, x, s, i)
    print("host dry-run ok")
