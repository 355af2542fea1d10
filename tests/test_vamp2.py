"""vamp2.py -- the reference's damped direct-form VAMP: the numpy restatement against outputs of the
reference's own class (tests/golden/vamp2_*.npz, written by make_golden.py from /root/reference/vamp2.py), and the CUDA kernel
(csrc/vamp2.cu, through ampsm_vamp2_detect) against both.  The method does not meet its exit test on these fixtures -- every
frame runs the 20 layers in the reference too -- so the whole 20-step trajectory is compared, at float32-rounding tolerances."""
import numpy as np
import pytest
import torch

import amp_sparc_spatialmodulation_b200 as pkg
from conftest import config_from_meta, load_golden
from oracle import amp_oracle as ao
from parity_utils import assert_counts_equal, counters_for, decision_mismatch_frames

DEV = "cuda:0"
NAMES = ["vamp2_d100", "vamp2_d097"]


@pytest.mark.parametrize("name", NAMES)
def test_vamp2_oracle_matches_reference(name):
    g = load_golden(name)
    cfg = config_from_meta(g["meta"])
    r = ao.vamp2_detect(g["U"], g["s"], g["Vh"], g["y"], g["sigma2"], np.asarray(cfg.symbols), cfg.L, cfg.M, cfg.N_Layers,
                        damping=g["meta"]["damping"], x_true=g["x"])
    assert (r["iters"] == g["iters"]).all()
    # gamma and mean var over all 20 iterations to 1e-5 relative (measured 7e-7), estimates to 1e-4 absolute (measured 1e-5)
    assert np.abs(r["traj"]["gamma"].T - g["gamma"]).max() <= 1e-5 * np.abs(g["gamma"]).max()
    assert np.abs(r["traj"]["var"].T - g["varm"]).max() <= 1e-5 * np.abs(g["varm"]).max()
    assert np.abs(r["traj"]["mse"].T - g["mse"]).max() <= 1e-6
    assert np.abs(r["xmmse"] - g["xmmse"]).max() < 1e-5 and np.abs(r["xmap"] - g["xmap"]).max() < 1e-4
    assert decision_mismatch_frames(cfg, r["xmap"], g["xmap"]).size == 0


@pytest.mark.gpu
@pytest.mark.parametrize("exp", ["f64", "f32"])
@pytest.mark.parametrize("name", NAMES)
def test_vamp2_kernel_matches_reference_goldens(name, exp):
    g = load_golden(name)
    F, N = g["x"].shape
    gam, varm, mse = np.zeros((F, 20)), np.zeros((F, 20)), np.zeros((F, 20))
    xmmse, xmap, iters = np.zeros((F, N), np.complex64), np.zeros((F, N), np.complex64), np.zeros(F, np.int32)
    for f in range(F):                      # per-frame factors with their own sigma2: one reference call per frame
        cfg = config_from_meta(g["meta"], batch=1, device=DEV)
        amp = pkg.VAMP2(cfg, g["meta"]["damping"], trajectory=True, exp=exp, shift="reference" if exp == "f64" else "section")
        snr = (cfg.Na / cfg.Nr) / float(g["sigma2"][f])
        t = lambda a: torch.as_tensor(np.ascontiguousarray(a), device=DEV)     # noqa: E731
        loss = amp(t(g["U"][f]), t(g["s"][f]), t(g["Vh"][f]), t(g["y"][f]).reshape(1, -1, 1), snr, t(g["x"][f]).reshape(1, -1, 1),
                   g["sym"][f], g["idx"][f])
        d = amp.last
        tr = d.traj.cpu().numpy()[0]
        gam[f], varm[f], mse[f] = tr[:, 0], tr[:, 1], tr[:, 2]
        xmmse[f], xmap[f], iters[f] = d.xmmse.cpu().numpy().ravel(), d.xmap.cpu().numpy().ravel(), int(d.iters.cpu()[0])
        want = counters_for(cfg, g["xmap"][f:f + 1], g["xmmse"][f:f + 1], g["x"][f:f + 1], g["sym"][f], g["idx"][f])
        assert_counts_equal(f"{name}[{f}]", d.counters_dict(), want)
        assert loss.loss['T'] == int(g["iters"][f])
    assert (iters == g["iters"]).all()
    tol = 2e-5 if exp == "f64" else 2e-4            # float32 exp: 2^-22 per exponential, not amplified (the iteration contracts)
    assert np.abs(gam - g["gamma"]).max() <= tol * np.abs(g["gamma"]).max()
    assert np.abs(varm - g["varm"]).max() <= tol * np.abs(g["varm"]).max()
    assert np.abs(mse - g["mse"]).max() <= 1e-5
    assert np.abs(xmmse - g["xmmse"]).max() < 1e-4 and np.abs(xmap - g["xmap"]).max() < 1e-3


@pytest.mark.gpu
def test_vamp2_batched_call_equals_oracle_and_per_frame_calls():
    """512 frames with per-frame factors in one call against the oracle on the same inputs: every counter, every exit iteration;
    shared factors (stride 0) equal per-frame copies bit for bit."""
    F = 512
    cfg = pkg.Config(32, 2, 16, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform', device=DEV)
    st = pkg.FrameStream(cfg, seed=17)
    snr = 10 ** 0.8
    H, y, x, lab, idx = st.frames(0, F, snr)
    U, s, Vh = torch.linalg.svd(H.cpu(), full_matrices=False)
    det = pkg.VAMP2(cfg, 0.97, outputs=True, exp='f64', shift='reference').detect(U, s, Vh, y, snr, x, lab, idx)
    ref = ao.vamp2_detect(U.numpy(), s.numpy(), Vh.numpy(), y.cpu().numpy(), (cfg.Na / cfg.Nr) / snr, np.asarray(cfg.symbols), cfg.L, cfg.M, 20,
                          damping=0.97)
    assert (det.iters.cpu().numpy() == ref["iters"]).all()
    assert np.abs(det.xmmse.cpu().numpy().reshape(F, -1) - ref["xmmse"]).max() < 1e-4
    pos = idx.cpu().numpy()
    want = counters_for(cfg, ref["xmap"], ref["xmmse"], x.cpu().numpy(), lab.cpu().numpy(), pos)
    assert_counts_equal("vamp2 batched", det.counters_dict(), want)
    sh = pkg.VAMP2(cfg, 0.97, outputs=True).detect(U[0], s[0], Vh[0], y[:32], snr, x[:32], lab[:32 * cfg.L], idx[:32 * cfg.L])
    pf = pkg.VAMP2(cfg, 0.97, outputs=True).detect(U[:1].expand(32, -1, -1).contiguous(), s[:1].expand(32, -1).contiguous(),
                                                   Vh[:1].expand(32, -1, -1).contiguous(), y[:32], snr, x[:32], lab[:32 * cfg.L], idx[:32 * cfg.L])
    assert torch.equal(sh.xmmse, pf.xmmse) and torch.equal(sh.iters, pf.iters)
