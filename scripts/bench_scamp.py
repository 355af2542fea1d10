"""SCAMP at the C4 instance: Config(512, 8, 32, 32, 3, 'tail', QPSK) -> A 1088 x 16384 shared by B
frames.  Prints device time per call, frame-iterations/s and the non-zero-block GEMM rate."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--ebn0-db", type=float, default=6.0)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--fixed", action="store_true")
a = ap.parse_args()
dev = "cuda:0"
cfg = pkg.Config(512, 8, 32, 32, 3, batch=a.frames, generator_mode='sparc', iterations=20, alphabet='QPSK',
                 channel_profile='uniform', channel_truncation='tail', device=dev)
np.random.seed(0)
torch.manual_seed(0)
ch, da = pkg.Channel(cfg), pkg.Data(cfg)
W, A = ch.generate_as_sparc()
x, sym, idx = da.generate_message()
snr_db = a.ebn0_db + 10 * np.log10(cfg.code_rate)
snr = 10 ** (snr_db / 10)
y = A @ x + ch.awgn(snr)
amp = pkg.SCAMP(cfg, outputs=False, early_exit=not a.fixed)
for _ in range(2):
    det = amp.detect(W, A, y, snr, x, sym, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.reps):
    det = amp.detect(W, A, y, snr, x, sym, idx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
c = det.counters_dict()
nnz = int((A != 0).sum())
flop = 2 * 8 * nnz * c["iters"]            # two complex GEMMs per iteration over the non-zero entries
print(f"SCAMP n={cfg.n} N={cfg.N} frames={a.frames} density={nnz / A.numel():.3f} mean T={c['iters'] / c['frames']:.2f} "
      f"fer={c['frame_err'] / c['frames']:.3f} ms/call={ms:.2f} frame-iter/s={c['iters'] / ms * 1e3:.3e} "
      f"nonzero-GEMM TFLOP/s={flop / ms / 1e9:.2f}")
