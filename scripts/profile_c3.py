"""Small fixed workload for ncu: VAMP 128 x 64 Na = 4 QPSK (BASELINE config 3) through the four-warps-per-frame kernel,
`--frames` frames sharing one SVD (one reference channel draw), a few launches."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=148 * 2 * 16)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--fixed", action="store_true")
ap.add_argument("--snr-db", type=float, default=2.0)
ap.add_argument("--double", action="store_true", help="complex128 factors: the register-resident DFMA kernel (csrc/vamp_dbl.cu)")
a = ap.parse_args()
cfg = pkg.Config(128, 4, 64, 1, 1, batch=a.frames, generator_mode='sparc', iterations=20, alphabet='QPSK',
                 channel_profile='uniform', device="cuda:0")
np.random.seed(0)
torch.manual_seed(0)
ch, da = pkg.Channel(cfg), pkg.Data(cfg)
snr = 10 ** (a.snr_db / 10)
x, sym, idx = da.generate_message()
_, A = ch.generate_as_sparc()
y = A @ x + ch.awgn(snr)
U, s, Vh = torch.linalg.svd(A, full_matrices=False)
if a.double:
    U, s, Vh, y = U.to(torch.complex128), s.to(torch.float64), Vh.to(torch.complex128), y.to(torch.complex128)
amp = pkg.VAMP(cfg, outputs=False, early_exit=not a.fixed)
for _ in range(a.launches):
    det = amp.detect(U, s, Vh, y, snr, x, sym, idx)
torch.cuda.synchronize()
c = det.counters_dict()
print("frames", c["frames"], "mean T", c["iters"] / c["frames"], "ier", c["index_err"] / (4 * c["frames"]))
