"""How many frames does the register-resident VAMP kernel decide differently from the generic float64-exponent kernel,
frame by frame (not net counts), next to the generic kernel's own float32-exp mode?  VAMP amplifies float32 rounding on
slowly converging frames, so net error counts of two evaluation orders differ by chance; this script shows the size of
that effect and whether one kernel is systematically worse."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from test_gpu_scale import c2, make_frames, svd_factors  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for seed in (21, 22, 23):
    cfg = c2(F, alphabet='16QAM', Na=1)
    H, y, x, lab, idx = make_frames(cfg, F, 12.0, seed=seed)
    U, s, Vh = svd_factors(H)
    snr = 10 ** 1.2
    sym = torch.as_tensor(cfg.symbols).to(H.device, torch.complex128)
    truth = x.abs().argmax(dim=1)

    def decided(det):
        xm = det.xmap.reshape(F, -1).to(torch.complex128)
        met = (xm.unsqueeze(-1) * sym.conj()).real.amax(dim=-1)
        return met.argmax(dim=1)
    out = {}
    for name, kw in (("fast", dict(kernel='fast')), ("gen64", dict(kernel='generic', exp='f64')), ("gen32", dict(kernel='generic', exp='f32'))):
        det = pkg.VAMP(cfg, outputs=True, **kw).detect(U, s, Vh, y, snr, x, lab, idx)
        out[name] = (decided(det), det.counters_dict()["index_err"], det.iters.cpu().numpy())
    d64 = out["gen64"][0]
    for name in ("fast", "gen32"):
        d = out[name][0]
        diff = d != d64
        better = ((d == truth) & (d64 != truth)).sum().item()
        worse = ((d != truth) & (d64 == truth)).sum().item()
        print(f"seed {seed} {name:>5s}: index_err {out[name][1]} (gen64 {out['gen64'][1]}), frames decided differently from gen64: "
              f"{int(diff.sum())} (right where gen64 is wrong: {better}, wrong where gen64 is right: {worse}), "
              f"mean T {out[name][2].mean():.3f} vs {out['gen64'][2].mean():.3f}", flush=True)
