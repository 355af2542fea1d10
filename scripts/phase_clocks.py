"""Phase timing of bamp_fast_kernel (development build, scripts/build_clk.sh): cycles per phase, per warp-iteration and per
frame, summed by lane 0 of every warp.  `AMPSM_LIB=.../libampsm_b200_clk.so python scripts/phase_clocks.py [--fixed]`.
Also prints the device time per launch, so the same script serves the occupancy experiments (AMPSM_CTAS_PER_SM=k)."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200 import _cabi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=148 * 8 * 128)
ap.add_argument("--fixed", action="store_true")
ap.add_argument("--iters", type=int, default=bench.ITERS)
ap.add_argument("--snr-db", type=float, default=15.0)
ap.add_argument("--vamp", action="store_true", help="time the VAMP fast kernel (per-frame SVD factors) instead")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = pkg.Config(bench.NT, bench.NA, bench.NR, 1, 1, batch=a.frames, generator_mode='sparc', iterations=a.iters,
                 alphabet=bench.ALPHABET, channel_profile='uniform', device="cuda:0")
H, y, x, labels, idx = bench.make_gpu_inputs(torch, cfg, a.frames, a.snr_db, dev, 1234)
lib = _cabi.lib()
have_clk = hasattr(lib, "ampsm_debug_clocks")
clk_fn = getattr(lib, "ampsm_debug_clocks_vamp" if a.vamp else "ampsm_debug_clocks", None)
snr = 10 ** (a.snr_db / 10)
if a.vamp:
    vamp = pkg.VAMP(cfg, kernel='auto', outputs=False, early_exit=not a.fixed)
    from amp_sparc_spatialmodulation_b200.vamp import svd_batched
    U, sv, Vh = svd_batched(H)                  # the library's Jacobi kernel (cuSOLVER's batched SVD takes minutes here)
    del H

    class _V:
        def detect(self, H_, y_, snr_, x_, lab_, idx_):
            return vamp.detect(U, sv, Vh, y_, snr_, x_, lab_, idx_)
    amp, H = _V(), None
else:
    amp = pkg.BAMP(cfg, kernel='auto', outputs=False, early_exit=not a.fixed)
for _ in range(2):
    det = amp.detect(H, y, snr, x, labels, idx)
torch.cuda.synchronize()
if have_clk:
    clk_fn(None, 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    det = amp.detect(H, y, snr, x, labels, idx)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
c = det.counters_dict()
fi = c["iters"]
print(f"{'VAMP' if a.vamp else 'BAMP'} ctas/sm={os.environ.get('AMPSM_CTAS_PER_SM', 'default')} fixed={a.fixed} iters={a.iters} frames={a.frames} mean T={fi / a.frames:.3f} "
      f"{ms:.3f} ms  {fi / ms * 1e3:.4e} frame-iter/s  {a.frames / ms * 1e3:.4e} frames/s", flush=True)
if have_clk:
    out = (C.c_ulonglong * 16)()
    clk_fn(out, 1)
    if a.vamp:
        names = ["row pass", "row reduce + LMMSE", "col pass + scalars", "col reduce + r", "denoiser", "Onsager + publish"]
        frame_phases = {6: "stage wait + y~", 7: "stage refill + init", 8: "epilogue"}
    else:
        names = ["row pass", "row reduce+publish", "col pass", "col reduce+xmap", "denoiser", "exit+publish"]
        frame_phases = {8: "tile load issue", 9: "outputs, counters", 10: "Loss input wait", 6: "Loss", 7: "prologue"}
    tot_it = fi * reps
    tot_fr = a.frames * reps
    s = 0.0
    for p in range(6):
        v = out[p] / tot_it
        s += v
        print(f"  {names[p]:>22s}: {v:8.1f} cycles / warp-iteration")
    print(f"  {'iteration total':>22s}: {s:8.1f}")
    for p, nm in frame_phases.items():
        print(f"  {nm:>22s}: {out[p] / tot_fr:8.1f} cycles / frame")
