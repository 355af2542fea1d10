"""BASELINE config 3 from the channel matrices: per-frame Jacobi SVD (one CTA per 64 x 128 matrix) + VAMP iterations in one call
(ampsm_vamp_detect_from_h), against the iterations alone on precomputed factors (scripts/time_c3.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200.simulate import device_frames  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 14
snr = 10 ** 0.2
cfg = pkg.Config(128, 4, 64, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform', device="cuda:0")
gen = torch.Generator(device="cuda:0").manual_seed(4321)
H, y, x, lab, idx = device_frames(cfg, F, snr, gen)
v = pkg.VAMP(cfg, outputs=False)
for _ in range(2):
    det = v.detect_from_channel(H, y, snr, x, lab, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
det = v.detect_from_channel(H, y, snr, x, lab, idx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
c = det.counters_dict()
print(f"C3 from channel: {F} frames {ms:.2f} ms  {F / ms * 1e3:.3e} frames/s  mean T={c['iters'] / F:.2f} ier={c['index_err'] / (4 * F):.5f}")
