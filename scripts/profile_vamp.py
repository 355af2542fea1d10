"""Small fixed workload for ncu: VAMP 64x32 16-QAM with per-frame factors, a few launches of the chosen kernel
(`--from-channel`: the Jacobi SVD kernel + the iterations)."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from test_gpu_scale import c2, make_frames, svd_factors  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=148 * 8 * 32)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--kernel", default="auto")
ap.add_argument("--fixed", action="store_true")
ap.add_argument("--from-channel", action="store_true")
ap.add_argument("--snr-db", type=float, default=15.0)
a = ap.parse_args()
cfg = c2(a.frames)
H, y, x, lab, idx = make_frames(cfg, a.frames, a.snr_db, seed=1234)
snr = 10 ** (a.snr_db / 10)
amp = pkg.VAMP(cfg, kernel=a.kernel, outputs=False, early_exit=not a.fixed)
if a.from_channel:
    for _ in range(a.launches):
        det = amp.detect_from_channel(H, y, snr, x, lab, idx)
else:
    U, s, Vh = svd_factors(H)
    for _ in range(a.launches):
        det = amp.detect(U, s, Vh, y, snr, x, lab, idx)
torch.cuda.synchronize()
c = det.counters_dict()
print("frames", c["frames"], "mean T", c["iters"] / c["frames"], "ier", c["index_err"] / c["frames"])
