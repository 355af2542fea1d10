import sys, os, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import amp_sparc_spatialmodulation_b200 as pkg
from test_gpu_scale import c2, make_frames
F = 262144
cfg = c2(F)
H, y, x, lab, idx = make_frames(cfg, F, 15.0, seed=5)
v = pkg.VAMP(cfg, outputs=False)
for _ in range(2): det = v.detect_from_channel(H, y, 10**1.5, x, lab, idx)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); det = v.detect_from_channel(H, y, 10**1.5, x, lab, idx); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1); c = det.counters_dict()
print(f"from_channel {F} frames {ms:.2f} ms {F/ms*1e3:.3e} frames/s ier {c['index_err']/F:.5f} mean T {c['iters']/F:.3f}")
from amp_sparc_spatialmodulation_b200.vamp import svd_batched
for _ in range(2): svd_batched(H)
torch.cuda.synchronize(); e0.record(); svd_batched(H); e1.record(); torch.cuda.synchronize()
print(f"svd_batched (U, s, Vh) {e0.elapsed_time(e1):.2f} ms {F/e0.elapsed_time(e1)*1e3:.3e} matrices/s")
