"""Tiny workloads for compute-sanitizer (racecheck / synccheck / memcheck): one detector call per case.

    compute-sanitizer --tool racecheck python scripts/sanitize_case.py --case bamp_c2
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200.simulate import device_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", required=True)
ap.add_argument("--frames", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
gen = torch.Generator(device=dev).manual_seed(3)


def cfg_of(Nt, Na, Nr, alphabet, F, Lin=1, Lh=1, trunc='trunc'):
    return pkg.Config(Nt, Na, Nr, Lin, Lh, batch=F, generator_mode='sparc', iterations=20, alphabet=alphabet,
                      channel_profile='uniform', channel_truncation=trunc, device=str(dev))


def svd(H):
    U, s, Vh = torch.linalg.svd(H.cpu(), full_matrices=False)
    return U.contiguous().to(dev), s.contiguous().to(dev), Vh.contiguous().to(dev)


case = a.case
if case in ("bamp_c1", "bamp_c2", "bamp_c2_na4"):
    shape = {"bamp_c1": (8, 1, 4, 'QPSK'), "bamp_c2": (64, 1, 32, '16QAM'), "bamp_c2_na4": (64, 4, 32, 'QPSK')}[case]
    F = a.frames or 600
    cfg = cfg_of(*shape, F)
    H, y, x, lab, idx = device_frames(cfg, F, 10 ** 1.2, gen)
    det = pkg.BAMP(cfg, kernel='fast', outputs=True).detect(H, y, 10 ** 1.2, x, lab, idx)
elif case in ("vamp_c2", "vamp_c3"):
    shape = {"vamp_c2": (64, 1, 32, '16QAM'), "vamp_c3": (128, 4, 64, 'QPSK')}[case]
    F = a.frames or 400
    cfg = cfg_of(*shape, F)
    snr = 10 ** (1.2 if case == "vamp_c2" else 0.2)
    H, y, x, lab, idx = device_frames(cfg, F, snr, gen)
    U, s, Vh = svd(H)
    det = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
elif case == "vamp_from_h":
    F = a.frames or 300
    cfg = cfg_of(64, 1, 32, '16QAM', F)
    H, y, x, lab, idx = device_frames(cfg, F, 10 ** 1.2, gen)
    det = pkg.VAMP(cfg, outputs=True).detect_from_channel(H, y, 10 ** 1.2, x, lab, idx)
elif case in ("scamp_tc", "scamp_simt"):
    F = a.frames or (160 if case == "scamp_tc" else 48)
    ccpu = pkg.Config(64, 2, 8, 8, 3, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                      channel_truncation='tail', device='cpu')
    np.random.seed(1)
    torch.manual_seed(1)
    ch, da = pkg.Channel(ccpu), pkg.Data(ccpu)
    W, A = ch.generate_as_sparc()
    x, lab, idx = da.generate_message()
    snr = 10 ** 0.6
    y = A @ x + ch.awgn(snr)
    cfg = cfg_of(64, 2, 8, 'QPSK', F, Lin=8, Lh=3, trunc='tail')
    det = pkg.SCAMP(cfg, outputs=True).detect(W, A, y, snr, x, lab, idx)
elif case == "bamp_generic":
    F = a.frames or 64
    cfg = cfg_of(24, 2, 12, 'QPSK', F)
    H, y, x, lab, idx = device_frames(cfg, F, 10.0, gen)
    det = pkg.BAMP(cfg, kernel='generic', outputs=True).detect(H, y, 10.0, x, lab, idx)
else:
    raise SystemExit(f"unknown case {case}")
torch.cuda.synchronize()
c = det.counters_dict()
print(case, "frames", c["frames"], "iters", c["iters"], "frame_err", c["frame_err"], "nan", c["nan_frames"])
