"""Frames per second of VAMP on 64 x 32 channels, generation included: (a) torch's generators + detect_from_channel (the
round-1 sweep path), (b) the generator kernel + detect_from_channel, (c) frames drawn inside the Jacobi SVD kernel
(ampsm_vamp_detect_generated: H never in HBM).  `--channel kronecker` is BASELINE config 5's workload."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200.simulate import device_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1 << 18)
ap.add_argument("--snr-db", type=float, default=6.0)
ap.add_argument("--channel", default="kronecker")
ap.add_argument("--rho", type=float, default=0.7)
ap.add_argument("--alphabet", default="QPSK")
a = ap.parse_args()
dev = torch.device("cuda:0")
F, snr = a.frames, 10 ** (a.snr_db / 10)
cfg = pkg.Config(64, 1, 32, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet=a.alphabet, channel_profile='uniform', device="cuda:0")
amp = pkg.VAMP(cfg, outputs=False)
st = pkg.FrameStream(cfg, seed=1, channel=a.channel, rho_t=a.rho, rho_r=a.rho)
gen = torch.Generator(device=dev).manual_seed(1)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def path_torch():
    H, y, x, lab, idx = device_frames(cfg, F, snr, gen, a.channel, a.rho, a.rho)
    return amp.detect_from_channel(H, y, snr, x, lab, idx)


def path_kernel():
    H, y, x, lab, idx = st.frames(0, F, snr)
    return amp.detect_from_channel(H, y, snr, x, lab, idx)


def path_fused():
    return amp.detect_generated(st, 0, F, snr)


def only_generate():
    return st.frames(0, F, snr)


for name, fn in (("torch generators + from_channel", path_torch), ("generator kernel + from_channel", path_kernel),
                 ("drawn inside the SVD kernel", path_fused)):
    ms, det = timed(fn)
    c = det.counters_dict()
    print(f"{name:>34s}: {ms:8.2f} ms  {F / ms * 1e3:.3e} frames/s  mean T={c['iters'] / F:.2f} fer={c['frame_err'] / F:.4f}", flush=True)
ms, _ = timed(only_generate)
print(f"{'generator kernel alone':>34s}: {ms:8.2f} ms  {F / ms * 1e3:.3e} frames/s  ({F * (32 * 64 + 32 + 64) * 8 / ms * 1e-6:.0f} GB/s written)")
