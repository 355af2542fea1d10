"""VAMP 64x32 diagnostics on the GPU: register-resident vs generic kernel (agreement and device time)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from test_gpu_scale import c2, make_frames, svd_factors, INT_KEYS  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
for alphabet, Na, snr_db in (("16QAM", 1, 12.0), ("QPSK", 1, 8.0), ("QPSK", 4, 6.0), ("QPSK", 4, 10.0)):
    cfg = c2(F, alphabet=alphabet, Na=Na)
    H, y, x, lab, idx = make_frames(cfg, F, snr_db, seed=21)
    U, s, Vh = svd_factors(H)
    snr = 10 ** (snr_db / 10)
    res = {}
    for name, kw in (("fast", dict(kernel='fast')), ("generic64", dict(kernel='generic', exp='f64')),
                     ("generic32", dict(kernel='generic', exp='f32'))):
        amp = pkg.VAMP(cfg, outputs=True, **kw)
        d = amp.detect(U, s, Vh, y, snr, x, lab, idx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        d = amp.detect(U, s, Vh, y, snr, x, lab, idx)
        e1.record()
        torch.cuda.synchronize()
        res[name] = (d, e0.elapsed_time(e1))
    a, b, c = res["fast"][0], res["generic64"][0], res["generic32"][0]
    for nm, u, v in (("fast-vs-g64", a, b), ("g32-vs-g64", c, b)):
        ia, ib = u.iters.cpu().numpy(), v.iters.cpu().numpy()
        cu, cv = u.counters_dict(), v.counters_dict()
        dd = (u.xmmse - v.xmmse).abs().reshape(F, -1).amax(dim=1)
        print(f"{alphabet} Na={Na} {snr_db}dB {nm}: iters eq {np.mean(ia == ib):.4f} |d|<=1 {np.mean(np.abs(ia - ib) <= 1):.4f} "
              f"|d|<=3 {np.mean(np.abs(ia - ib) <= 3):.4f} meanT {ia.mean():.2f}/{ib.mean():.2f} "
              f"xmmse med {float(dd.median()):.2e} q97 {float(torch.quantile(dd, 0.97)):.2e} "
              f"counter diffs { {k: (cu[k], cv[k]) for k in INT_KEYS if cu[k] != cv[k]} }")
    it = res["fast"][0].counters_dict()["iters"]
    print(f"   time ms fast {res['fast'][1]:.3f} generic64 {res['generic64'][1]:.3f} generic32 {res['generic32'][1]:.3f} "
          f"fast frame-iter/s {it / res['fast'][1] * 1e3:.3e}")
