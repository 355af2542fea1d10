"""BAMP on a structured ISI channel: block-convolution operator (ampsm_bamp_detect_taps) against the dense kernel.

    python scripts/bench_isi.py [frames]

Config: Nt=128, Na=4, Nr=24, Lin=32, Lh=3, 'tail', QPSK (the regime of the reference's published simulations,
SURVEY.md section 8f.1): n = 816, N = 4096, dense H = 26.7 MB per channel draw, taps = 73.7 KB.  One channel draw shared by
the frames of a call (the reference's layout), exit disabled (exactly 20 iterations per frame).
Algorithmic flop per frame-iteration (SURVEY.md 8d with the non-zero blocks only): 20 Lh Lin Nr Nt + 18 N K + 30 n + 20 N.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    dev = "cuda:0"
    cfg = pkg.Config(128, 4, 24, 32, 3, batch=frames, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', channel_truncation='tail', device=dev)
    np.random.seed(0)
    torch.manual_seed(0)
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    H, taps = ch.generate_channel(return_taps=True)
    x, sym, idx = da.generate_message()
    snr = 10 ** (6 / 10)
    y = H @ x + ch.awgn(snr)
    flop = 20 * cfg.Lh * cfg.Lin * cfg.Nr * cfg.Nt + 18 * cfg.N * cfg.K + 30 * cfg.n + 20 * cfg.N
    out = {}
    for tag, structured in (("taps", True), ("dense", False)):
        amp = pkg.BAMP(cfg, early_exit=False, outputs=False, structured=structured)
        run = (lambda: amp.detect_taps(taps, y, snr, x, sym, idx)) if structured else (lambda: amp.detect(H, y, snr, x, sym, idx))
        ms = timed(run)
        c = run().counters_dict()
        out[tag] = dict(ms=ms, frame_iter_per_s=frames * 20 / ms * 1e3, tflops=frames * 20 * flop / ms * 1e-9,
                        ier=c["index_err"] / (frames * cfg.L), nan=c["nan_frames"])
        print(tag, out[tag], flush=True)
    print("speed-up taps/dense: %.2fx; algorithmic flop per frame-iteration %d" % (out["dense"]["ms"] / out["taps"]["ms"], flop))


if __name__ == "__main__":
    main()
