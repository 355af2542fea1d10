"""Small fixed workload for ncu: BAMP 64x32 16-QAM (C2), `--frames` frames, a few launches of the chosen kernel."""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=148 * 8 * 32)
ap.add_argument("--launches", type=int, default=3)
ap.add_argument("--kernel", default="auto")
ap.add_argument("--fixed", action="store_true")
ap.add_argument("--snr-db", type=float, default=15.0)
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = pkg.Config(bench.NT, bench.NA, bench.NR, 1, 1, batch=a.frames, generator_mode='sparc', iterations=bench.ITERS,
                 alphabet=bench.ALPHABET, channel_profile='uniform', device="cuda:0")
H, y, x, labels, idx = bench.make_gpu_inputs(torch, cfg, a.frames, a.snr_db, dev, 1234)
amp = pkg.BAMP(cfg, kernel=a.kernel, outputs=False, early_exit=not a.fixed)
for _ in range(a.launches):
    det = amp.detect(H, y, 10 ** (a.snr_db / 10), x, labels, idx)
torch.cuda.synchronize()
c = det.counters_dict()
print("frames", c["frames"], "mean T", c["iters"] / c["frames"], "fer", c["frame_err"] / c["frames"])
