"""Device-time throughput of the BASELINE.json configurations that bench.py does not headline: C1 (BAMP 8 x 4 QPSK) and
C3 (VAMP 128 x 64, Na = 4, QPSK; complex64 and complex128 factors), early exit as the reference.  Frames are generated on
the device; the SVD of C3 is taken once by torch (cuSOLVER) and shared by the frames of the call, like one reference
channel draw with `res` frames -- per-frame factors would need 64 KiB x frames of Vh."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200.simulate import device_frames  # noqa: E402

DEV = "cuda:0"


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def c1(frames=1 << 20, snr_db=10.0):
    cfg = pkg.Config(8, 1, 4, 1, 1, batch=frames, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', device=DEV)
    gen = torch.Generator(device=DEV).manual_seed(1)
    snr = 10 ** (snr_db / 10)
    H, y, x, lab, idx = device_frames(cfg, frames, snr, gen)
    amp = pkg.BAMP(cfg, outputs=False)
    ms, det = timed(lambda: amp.detect(H, y, snr, x, lab, idx))
    c = det.counters_dict()
    print(f"C1 BAMP 8x4 QPSK {snr_db} dB: {frames} frames, mean T={c['iters'] / frames:.2f}, {ms:.3f} ms, "
          f"{c['iters'] / ms * 1e3:.3e} frame-iter/s, {frames / ms * 1e3:.3e} frames/s, "
          f"{frames * 352 / ms * 1e-6:.0f} GB/s algorithmic, ier={c['index_err'] / frames:.4f}", flush=True)


def c3(frames, double, snr_db=2.0, per_frame=False):
    cfg = pkg.Config(128, 4, 64, 1, 1, batch=frames, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', device=DEV)
    np.random.seed(0)
    torch.manual_seed(0)
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    snr = 10 ** (snr_db / 10)
    x, sym, idx = da.generate_message()
    if per_frame:
        A = torch.stack([ch.generate_as_sparc()[1] for _ in range(frames)])
        y = A @ x + ch.awgn(snr)
    else:
        _, A = ch.generate_as_sparc()
        y = A @ x + ch.awgn(snr)
    U, s, Vh = torch.linalg.svd(A, full_matrices=False)
    if double:
        U, s, Vh, y = U.to(torch.complex128), s.to(torch.float64), Vh.to(torch.complex128), y.to(torch.complex128)
    amp = pkg.VAMP(cfg, outputs=False)
    ms, det = timed(lambda: amp.detect(U, s, Vh, y, snr, x, sym, idx), reps=3)
    c = det.counters_dict()
    flop = 16 * 64 * 128 + 18 * 128 * 4 + 40 * 128 + 10 * 64
    print(f"C3 VAMP 128x64 Na=4 QPSK {'c128' if double else 'c64'} {'per-frame' if per_frame else 'shared'} factors {snr_db} dB: "
          f"{frames} frames, mean T={c['iters'] / frames:.2f}, {ms:.3f} ms, {c['iters'] / ms * 1e3:.3e} frame-iter/s, "
          f"{c['iters'] * flop / ms * 1e-9:.2f} TFLOP/s algorithmic, ier={c['index_err'] / (frames * 4):.4f}", flush=True)


if __name__ == "__main__":
    c1()
    c3(16384, False)
    c3(4096, False, per_frame=True)
    c3(16384, True)
