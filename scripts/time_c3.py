"""Device time of BASELINE config 3 (VAMP 128 x 64, Na = 4, QPSK) through the four-warps-per-frame kernel with per-frame
factors resident in HBM -- the same workload as bench.py's `vamp_c3` leg, for kernel experiments.  With the development
build (scripts/build_clk.sh, AMPSM_LIB=.../libampsm_b200_clk.so) it also prints the kernel's phase clocks."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import amp_sparc_spatialmodulation_b200 as pkg  # noqa: E402
from amp_sparc_spatialmodulation_b200 import _cabi  # noqa: E402
from amp_sparc_spatialmodulation_b200.simulate import device_frames  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1 << 14)
ap.add_argument("--snr-db", type=float, default=2.0)
ap.add_argument("--na", type=int, default=4)
ap.add_argument("--alphabet", default="QPSK")
a = ap.parse_args()
dev = torch.device("cuda:0")
F = a.frames
cfg = pkg.Config(128, a.na, 64, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet=a.alphabet,
                 channel_profile='uniform', device="cuda:0")
snr = 10 ** (a.snr_db / 10)
gen = torch.Generator(device=dev).manual_seed(4321)
H, y, x, lab, idx = device_frames(cfg, F, snr, gen)
Us, ss, Vs = [], [], []
for H1 in H.split(4096):                                    # thin SVD through the Hermitian eigenproblem of H H^H (float64)
    Hd = H1.to(torch.complex128)
    wv, V = torch.linalg.eigh(Hd @ Hd.mH)
    wv, V = wv.flip(-1), V.flip(-1)
    sv = wv.clamp_min(0).sqrt()
    Us.append(V.to(torch.complex64)), ss.append(sv.to(torch.float32))
    Vs.append(((V.mH @ Hd) / sv.unsqueeze(-1)).to(torch.complex64))
U, s, Vh = torch.cat(Us).contiguous(), torch.cat(ss).contiguous(), torch.cat(Vs).contiguous()
del H, Us, ss, Vs
lib = _cabi.lib()
clk_fn = getattr(lib, "ampsm_debug_clocks_quad", None) if hasattr(lib, "ampsm_debug_clocks_quad") else None
flop = 16 * 64 * 128 + 18 * 128 * 4 + 40 * 128 + 10 * 64
for tag, ee in (("exit", True), ("fixed_T", False)):
    v = pkg.VAMP(cfg, outputs=False, early_exit=ee)
    for _ in range(2):
        v.detect(U, s, Vh, y, snr, x, lab, idx)
    torch.cuda.synchronize()
    if clk_fn:
        clk_fn(None, 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        det = v.detect(U, s, Vh, y, snr, x, lab, idx)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    c = det.counters_dict()
    fi = c["iters"]
    print(f"C3 {tag}: frames={F} mean T={fi / F:.3f} {ms:.3f} ms  {fi / ms * 1e3:.4e} frame-iter/s  {fi * flop / ms * 1e-9:.2f} TFLOP/s  "
          f"ier={c['index_err'] / (a.na * F):.5f} nan={c['nan_frames']}", flush=True)
    if clk_fn:
        out = (C.c_ulonglong * 16)()
        clk_fn(out, 1)
        names = ["scalars + row pass", "barrier 1", "row reduce + LMMSE", "column pass", "col reduce + denoiser", "var sum + barrier 2 + r~"]
        tot = fi * reps * 4                   # warp-iterations
        ssum = 0.0
        for p in range(6):
            vv = out[p] / tot
            ssum += vv
            print(f"  {names[p]:>26s}: {vv:8.1f} cycles / warp-iteration")
        print(f"  {'iteration total':>26s}: {ssum:8.1f}")
        for p, nm in ((6, "frame prologue"), (7, "frame epilogue")):
            print(f"  {nm:>26s}: {out[p] / (F * reps * 4):8.1f} cycles / warp-frame")
