"""GPU parity: the CUDA path (through the C-ABI, via the Python mirror) against the golden fixtures (outputs of the
reference itself) and against the numpy oracle on the same inputs.  Run on the B200 box: pytest -m gpu."""
import numpy as np
import pytest
import torch

import amp_sparc_spatialmodulation_b200 as pkg
from amp_sparc_spatialmodulation_b200 import _cabi
from conftest import config_from_meta, load_golden
from oracle import amp_oracle as ao
from oracle import loss_oracle as lo
from parity_utils import INT_KEYS, assert_counts_equal, check_trajectory, counters_for, decision_mismatch_frames

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def t(a):
    return torch.as_tensor(np.ascontiguousarray(a)).to(DEV)


def global_idx(g, N):
    F = g["x"].shape[0]
    return (g["idx"].reshape(F, -1) + (np.arange(F) * N)[:, None]).reshape(-1)


def run_bamp_golden(name, per_snr=True, **kw):
    """Run every frame of a BAMP fixture (grouped by SNR point: one call per sigma2)."""
    g = load_golden(name)
    F = g["x"].shape[0]
    N = g["x"].shape[1]
    out = dict(xmap=np.zeros((F, N), np.complex64), xmmse=np.zeros((F, N), np.complex64), iters=np.zeros(F, np.int32),
               traj=np.zeros((F, 20, 3), np.float32), counters=[])
    for snr_db in sorted(set(g["snr_db"].tolist())):
        sel = np.nonzero(g["snr_db"] == snr_db)[0]
        cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
        amp = pkg.BAMP(cfg, trajectory=True, **kw)
        idx = (g["idx"][sel].reshape(len(sel), -1) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
        amp(t(g["H"][sel]), t(g["y"][sel]).unsqueeze(-1), 10 ** (snr_db / 10), t(g["x"][sel]).unsqueeze(-1),
            g["sym"][sel].reshape(-1), idx)
        d = amp.last
        out["xmap"][sel] = d.xmap.cpu().numpy().reshape(len(sel), N)
        out["xmmse"][sel] = d.xmmse.cpu().numpy().reshape(len(sel), N)
        out["iters"][sel] = d.iters.cpu().numpy()
        out["traj"][sel] = d.traj.cpu().numpy()
        want = counters_for(cfg, g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel], idx)
        out["counters"].append((snr_db, d.counters_dict(), want, amp.L))
    return g, out


@pytest.mark.parametrize("name", ["bamp_c1", "bamp_c2", "bamp_isi", "bamp_seg"])
@pytest.mark.parametrize("mode", [dict(kernel="generic", exp="f64", shift="reference"),
                                  dict(kernel="generic", exp="f32", shift="section"),
                                  dict(kernel="auto", exp="f32", shift="section"),
                                  dict(kernel="fast", exp="f32", shift="section")])
def test_bamp_matches_reference_goldens(name, mode):
    if mode["kernel"] == "fast" and name != "bamp_c2":
        pytest.skip("'auto' already takes the one-warp kernel for this shape (or none fits)")
    g, out = run_bamp_golden(name, **mode)
    cfg = config_from_meta(g["meta"])
    assert np.abs(out["iters"] - g["iters"]).max() <= 1, (out["iters"], g["iters"])
    assert (out["iters"] == g["iters"]).mean() >= 0.85
    # a non-converging frame amplifies rounding (the oracle itself is 7e-4 away from the reference on bamp_seg)
    per_frame = np.abs(out["xmmse"] - g["xmmse"]).max(axis=1)
    assert per_frame.max() < 1e-2 and np.median(per_frame) < 1e-4
    check_trajectory(name + ".tau", out["traj"][:, :, 0], g["tau"])
    check_trajectory(name + ".var", out["traj"][:, :, 1], g["varm"])
    # float32 exp: the tiny estimates of the inactive antennas are differences of nearly equal exponentials, good to
    # ~2^-22/|q| only; an MSE below 1e-7 of the unit symbol power (-70 dB) is therefore compared in float64 mode only
    check_trajectory(name + ".mse", out["traj"][:, :, 2], g["mse"], loose=0.2, floor=1e-9 if mode["exp"] == "f64" else 1e-7)
    # hard decisions and every error count identical to the reference's own Loss on its own estimates
    assert decision_mismatch_frames(cfg, out["xmap"], g["xmap"]).size == 0
    for snr_db, have, want, _ in out["counters"]:
        assert_counts_equal(f"{name}@{snr_db}dB", have, want)
        assert have["sqerr"] == pytest.approx(want["sqerr"], rel=2e-3, abs=1e-9)


def test_bamp_loss_dict_matches_reference_batch_loss():
    """One call over all frames of the C1 fixture at one SNR: the 14 rates equal the reference Loss with B=frames."""
    g = load_golden("bamp_c1")
    sel = np.nonzero(g["snr_db"] == 10)[0]
    N = g["x"].shape[1]
    cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
    idx = (g["idx"][sel].reshape(len(sel), -1) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
    L = pkg.BAMP(cfg)(t(g["H"][sel]), t(g["y"][sel]).unsqueeze(-1), 10.0, t(g["x"][sel]).unsqueeze(-1),
                      g["sym"][sel].reshape(-1), idx)
    want = counters_for(cfg, g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel], idx)
    ref = lo.rates_from_counters(want, dict(Na=cfg.Na, Lin=cfg.Lin), cfg.index_bits, cfg.symbol_bits)
    for k in L.keys:
        if np.isnan(ref[k]):
            continue
        assert float(L.loss[k]) == pytest.approx(ref[k], rel=2e-3, abs=1e-9), k
    assert L.loss['T'] == pytest.approx(g["iters"][sel].mean(), abs=0.1)


@pytest.mark.parametrize("name,double", [("vamp_c3", False), ("vamp_isi", False), ("vamp_c2", False), ("vamp_c2_na4", False),
                                         ("vamp_c5_rho07", False), ("vamp_c5_rho09", False), ("vamp_c3_c128", True)])
@pytest.mark.parametrize("exp", ["f64", "f32", "f32-generic"])
def test_vamp_matches_reference_goldens(name, double, exp):
    """exp = 'f32' lets the library choose: the register-resident kernels for the 64 x 32 fixtures (vamp_c2*, one warp
    per frame) and for the 128 x 64 one (vamp_c3, four warps per frame), the generic one elsewhere; 'f32-generic' pins the
    generic kernel on the same fixtures."""
    kernel = "auto"
    if exp == "f32-generic":
        if not (name.startswith("vamp_c2") or name == "vamp_c3"):
            pytest.skip("'f32' already runs the generic kernel for this shape")
        exp, kernel = "f32", "generic"
    if double and exp == "f32":
        pytest.skip("complex128 path always uses float64 exponents")
    g = load_golden(name)
    F, N = g["x"].shape
    ct = torch.complex128 if double else torch.complex64
    s2t = np.zeros((F, 20))
    varm = np.zeros((F, 20))
    xmmse = np.zeros((F, N), np.complex64)
    xmap = np.zeros((F, N), np.complex128)
    iters = np.zeros(F, np.int32)
    slow = []
    for f in range(F):          # per-frame factors with their own sigma2: one call per frame
        cfg = config_from_meta(g["meta"], batch=1, device=DEV)
        amp = pkg.VAMP(cfg, trajectory=True, exp=exp, shift="reference" if exp == "f64" else "section", kernel=kernel)
        snr = (cfg.Na / cfg.Nr) / float(g["sigma2"][f])
        amp(t(g["U"][f]).to(ct), t(g["s"][f]).to(torch.float64 if double else torch.float32), t(g["Vh"][f]).to(ct),
            t(g["y"][f]).to(ct).reshape(1, -1, 1), snr, t(g["x"][f]).reshape(1, -1, 1), g["sym"][f], g["idx"][f])
        d = amp.last
        tr = d.traj.cpu().numpy()[0]
        s2t[f], varm[f] = tr[:, 0], tr[:, 1]
        xmmse[f] = d.xmmse.cpu().numpy().ravel()
        xmap[f] = d.xmap.cpu().numpy().ravel()
        iters[f] = int(d.iters.cpu()[0])
        want = counters_for(cfg, g["xmap"][f:f + 1], g["xmmse"][f:f + 1], g["x"][f:f + 1], g["sym"][f], g["idx"][f])
        # a frame the reference itself could not bring to the exit test wanders chaotically in float32: its final
        # estimate (and decision) is not comparable between two evaluation orders; every converged frame must count alike
        # (slow = more than half of the iteration budget: vamp_c2 frame 1 takes 16 iterations in the reference, 11 in the oracle)
        if int(g["iters"][f]) <= cfg.N_Layers // 2:
            assert_counts_equal(f"{name}[{f}]", d.counters_dict(), want)
        else:
            slow.append((f, int(g["iters"][f]), iters[f]))
    # the frames left out of the count comparison are listed and their share is bounded (half of the fixture at most: vamp_c3 sits at
    # 2 dB, where three of its six frames use the whole iteration budget in the reference itself)
    print(f"{name} [{kernel}/{exp}]: frames outside the count comparison (frame, reference iterations, kernel iterations): {slow}")
    assert len(slow) <= max(1, F // 2), slow
    tight = 5e-7 if double else 1e-4     # see tests/test_oracle_golden.py for why not 1e-10
    if name.startswith("vamp_c5"):
        tight = 1e-3        # sigma2_tilde is posterior tail mass from the first iteration on (tests/test_oracle_golden.py)
    # (vamp_c2: sigma2_tilde is tail mass from iteration 2 on -- see tests/test_oracle_golden.py)
    for it in range(1 if name == "vamp_c2" else 2):
        assert np.abs(s2t[:, it] - g["sigma2t"][:, it]).max() <= max(tight, 2e-7) * np.abs(g["sigma2t"][:, it]).max() + 1e-12
        assert np.abs(varm[:, it] - g["varm"][:, it]).max() <= max(tight, 2e-7) * np.abs(g["varm"][:, it]).max() + 1e-12
    if double:
        assert (iters == g["iters"]).all()
        assert np.abs(xmmse - g["xmmse"]).max() < 1e-6
        assert np.abs(xmap - g["xmap"]).max() < 1e-6
    else:
        # beyond iteration 2 only a two-sided bound is meaningful: compare with the oracle's
        # own distance from the reference
        cfg = config_from_meta(g["meta"])
        r = ao.vamp_detect(g["U"], g["s"], g["Vh"], g["y"], g["sigma2"], cfg.Na / cfg.Nt, cfg.symbols, cfg.L, cfg.M, 20)
        d_oracle = np.abs(r["traj"]["sigma2"].T - g["sigma2t"]) / g["sigma2t"]
        d_kernel = np.abs(s2t - g["sigma2t"]) / g["sigma2t"]
        assert np.median(d_kernel) <= max(5e-2, 3 * np.median(d_oracle))
        solid = g["iters"] <= cfg.N_Layers // 2               # slowly / never converging frames are chaotic in float32
        assert (iters[g["iters"] <= 4] == g["iters"][g["iters"] <= 4]).all()
        per_frame = np.abs(xmmse - g["xmmse"]).max(axis=1)
        assert per_frame[solid].max(initial=0.0) < 2e-3 and solid.mean() > 0.4
    cfg = config_from_meta(g["meta"])
    conv = np.nonzero(g["iters"] <= cfg.N_Layers // 2)[0]
    assert decision_mismatch_frames(cfg, xmap[conv], g["xmap"][conv]).size == 0


@pytest.mark.parametrize("name", ["vamp_c5_rho07", "vamp_c5_rho09"])
def test_vamp_from_correlated_channel_matches_reference_goldens(name):
    """BASELINE config 5: the reference's VAMP fed torch.linalg.svd of a Kronecker-correlated channel (fixture written by
    tests/golden/make_golden.py) against ``detect_from_channel`` = in-kernel one-sided Jacobi SVD of the same H + iterations in
    one C-ABI call (vamp_model.py:56-61).  The two factorisations differ by unitary phases only, VAMP uses V f(S) V^H and
    V S U^H: estimates to float32 accuracy, identical exit iterations, decisions and error counts on every frame."""
    g = load_golden(name)
    F, N = g["x"].shape
    for snr_db in sorted(set(g["snr_db"].tolist())):
        sel = np.nonzero(g["snr_db"] == snr_db)[0]
        cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
        idx = (g["idx"][sel].reshape(len(sel), -1) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
        snr = 10 ** (snr_db / 10)
        d = pkg.VAMP(cfg, outputs=True).detect_from_channel(t(g["A"][sel]), t(g["y"][sel]), snr, t(g["x"][sel]), g["sym"][sel].reshape(-1), idx)
        xmmse = d.xmmse.cpu().numpy().reshape(len(sel), N)
        xmap = d.xmap.cpu().numpy().reshape(len(sel), N)
        iters = d.iters.cpu().numpy()
        assert np.abs(iters - g["iters"][sel]).max() <= 1 and (iters == g["iters"][sel]).mean() >= 0.8, (iters, g["iters"][sel])
        assert np.abs(xmmse - g["xmmse"][sel]).max() < 2e-4, np.abs(xmmse - g["xmmse"][sel]).max()
        assert decision_mismatch_frames(cfg, xmap, g["xmap"][sel].astype(np.complex64)).size == 0
        want = counters_for(cfg, g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel], idx)
        assert_counts_equal(f"{name}@{snr_db}dB", d.counters_dict(), want)


def test_vamp_batched_shared_factors_equals_per_frame_calls():
    """Frames sharing one SVD in a single call evolve exactly like separate batch=1 calls (per-frame pooled variance)."""
    g = load_golden("vamp_isi")
    cfg1 = config_from_meta(g["meta"], batch=1, device=DEV)
    cfgF = config_from_meta(g["meta"], batch=4, device=DEV)
    N = g["x"].shape[1]
    f0 = 8                                                     # frames 8.. share sigma2 (second SNR point)
    ys = t(np.stack([g["y"][f0 + i] for i in range(4)])).unsqueeze(-1)
    # feed all four observations through frame f0's factors: not a decode, just a determinism/equivalence check
    snr = (cfg1.Na / cfg1.Nr) / float(g["sigma2"][f0])
    d_big = pkg.VAMP(cfgF).detect(t(g["U"][f0]), t(g["s"][f0]), t(g["Vh"][f0]), ys, snr)
    for i in range(4):
        d_one = pkg.VAMP(cfg1).detect(t(g["U"][f0]), t(g["s"][f0]), t(g["Vh"][f0]), ys[i:i + 1], snr)
        assert torch.equal(d_big.xmmse[i], d_one.xmmse[0])
        assert int(d_big.iters[i]) == int(d_one.iters[0])


@pytest.mark.parametrize("structured", [False, 'auto'])
@pytest.mark.parametrize("exp", ["f64", "f32"])
def test_scamp_matches_reference_goldens(exp, structured):
    """structured=False: dense SIMT tiles (4 frames per matrix); 'auto': the reference's design matrices are block-Toeplitz, so
    the call runs the structured tensor-core kernels (tensor TMA + tcgen05, csrc/scamp_st.cu) from the taps."""
    g = load_golden("scamp_small")
    N = g["x"].shape[1]
    for ai in range(g["A"].shape[0]):
        sel = np.nonzero(g["a_of_frame"] == ai)[0]
        cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
        amp = pkg.SCAMP(cfg, trajectory=True, exp=exp, shift="reference" if exp == "f64" else "section", structured=structured)
        if structured:
            assert amp._taps_of(t(g["A"][ai])) is not None, "the reference's own design matrix must be recognised as structured"
        idx = (g["idx"][sel].reshape(len(sel), -1) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
        snr = (cfg.Na / cfg.Nr) / float(g["sigma2"][sel[0]])
        amp(t(g["W"][ai]), t(g["A"][ai]), t(g["y"][sel]).unsqueeze(-1), snr, t(g["x"][sel]).unsqueeze(-1),
            g["sym"][sel].reshape(-1), idx)
        d = amp.last
        iters = d.iters.cpu().numpy()
        assert np.abs(iters - g["iters"][sel]).max() <= 1, (iters, g["iters"][sel])
        xmmse = d.xmmse.cpu().numpy().reshape(len(sel), N)
        assert np.abs(xmmse - g["xmmse"][sel]).max() < 2e-3
        tr = d.traj.cpu().numpy()
        check_trajectory("scamp.tau", tr[:, :, 0], g["tau"][sel])
        check_trajectory("scamp.psi", tr[:, :, 1], g["psim"][sel])
        want = counters_for(cfg, g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel], idx)
        assert_counts_equal(f"scamp[{ai}]", d.counters_dict(), want)
        assert decision_mismatch_frames(cfg, d.xmap.cpu().numpy().reshape(len(sel), N), g["xmap"][sel]).size == 0


@pytest.mark.parametrize("name", ["loss_qpsk", "loss_16qam"])
def test_loss_kernel_matches_reference_loss(name):
    g = load_golden(name)
    F = g["x"].shape[0]
    cfg = config_from_meta(g["meta"], batch=F, device=DEV)
    L = pkg.Loss(cfg)
    L(t(g["xmap"]).unsqueeze(-1), t(g["xmmse"]).unsqueeze(-1), t(g["x"]).unsqueeze(-1), g["sym"], g["idx"], 3)
    for k, want in zip(L.keys, g["loss"]):
        tol = 1e-5 if k.startswith("nMSE") else 1e-12
        assert float(L.loss[k]) == pytest.approx(want, abs=tol), k
    want = counters_for(cfg, g["xmap"], g["xmmse"], g["x"], g["sym"], g["idx"])
    assert_counts_equal(name, L.counters, want)


def test_loss_kernel_nan_estimates_follow_argmax_rule():
    """NaN in xmap: np.argmax picks the first NaN (loss.py:296); the kernel must decide the same and count it."""
    g = load_golden("loss_qpsk")
    F = g["x"].shape[0]
    cfg = config_from_meta(g["meta"], batch=F, device=DEV)
    xmap = g["xmap"].copy()
    xmap[1, 5] = np.nan + 0j
    xmap[2, 3] = complex(0.1, np.nan)
    L = pkg.Loss(cfg)
    L(t(xmap).unsqueeze(-1), t(g["xmmse"]).unsqueeze(-1), t(g["x"]).unsqueeze(-1), g["sym"], g["idx"], 1)
    want = counters_for(cfg, xmap, g["xmmse"], g["x"], g["sym"], g["idx"])
    assert_counts_equal("nan", L.counters, want)
    assert L.counters["nan_frames"] == 2


@pytest.mark.parametrize("name", ["bamp_random", "bamp_random_isi"])
def test_bamp_random_mode_matches_reference_goldens(name):
    """generator_mode='random' (bamp.py:46,79-97; loss.py:252-280) through the generic kernel, one call per SNR point:
    exit iterations, estimates, trajectories and every error count against the reference's own run."""
    g = load_golden(name)
    F, N = g["x"].shape
    cfg1 = config_from_meta(g["meta"])
    per = cfg1.Lin * cfg1.Na                                    # labels per frame
    for snr_db in sorted(set(g["snr_db"].tolist())):
        sel = np.nonzero(g["snr_db"] == snr_db)[0]
        cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
        amp = pkg.BAMP(cfg, trajectory=True, exp="f64")
        idx = (g["idx"][sel].reshape(len(sel), per) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
        amp(t(g["H"][sel]), t(g["y"][sel]).unsqueeze(-1), 10 ** (snr_db / 10), t(g["x"][sel]).unsqueeze(-1),
            g["sym"][sel].reshape(-1), idx)
        d = amp.last
        iters = d.iters.cpu().numpy()
        assert np.abs(iters - g["iters"][sel]).max() <= 1
        conv = g["iters"][sel] <= cfg.N_Layers // 2
        per_frame = np.abs(d.xmmse.cpu().numpy().reshape(len(sel), N) - g["xmmse"][sel]).max(axis=1)
        assert per_frame[conv].max(initial=0.0) < 2e-3
        tr = d.traj.cpu().numpy()
        check_trajectory(name + ".tau", tr[:, :, 0], g["tau"][sel])
        check_trajectory(name + ".var", tr[:, :, 1], g["varm"][sel], loose=0.2)
        # Loss kernel alone on the REFERENCE's estimates: every count as the oracle (which reproduces the reference's rates)
        L = pkg.Loss(cfg)
        c = L._count(t(g["xmap"][sel]), t(g["xmmse"][sel]), t(g["x"][sel]), g["sym"][sel].reshape(-1), idx)
        want = lo.error_counters(g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel].reshape(-1), idx, cfg.symbols,
                                 cfg.gray, dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin), decision='random')
        assert_counts_equal(f"{name}@{snr_db}dB loss", c, want)
        assert c["sqerr"] == pytest.approx(want["sqerr"], rel=1e-5, abs=1e-9)
        # fused epilogue of the detector on its own estimates: converged frames decide like the reference
        have = d.counters_dict()
        assert have["frames"] == len(sel) and have["nan_frames"] == 0
        if conv.all():
            mine = lo.error_counters(g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel].reshape(-1), idx, cfg.symbols,
                                     cfg.gray, dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin), decision='random')
            assert_counts_equal(f"{name}@{snr_db}dB fused", have, mine)


def test_shrink_family_matches_reference_goldens():
    """Shrink 'bayes', 'shrinkOOK', sw_shrinkOOK (shrink.py:58-157) through ampsm_shrink: float32 arithmetic like the
    reference, tolerance 2e-6 absolute on outputs in [0, 1] (two float32 exp evaluations apart); a scalar cov
    broadcasts like the 0-dim gamma of vamp2.py:59; 'shrink' / 'lasso' fail as they do in the reference."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "shrink.npz"))
    cq = pkg.Config(16, 2, 8, 1, 1, batch=6, generator_mode='random', alphabet='QPSK', channel_profile='uniform', device=DEV)
    co = pkg.Config(16, 2, 8, 1, 1, batch=6, generator_mode='segmented', alphabet='OOK', channel_profile='uniform', device=DEV)
    rq, cvq = t(g["r_q"]).unsqueeze(-1), t(g["cov_q"]).unsqueeze(-1)
    ro, cvo = t(g["r_o"]).unsqueeze(-1), t(g["cov_o"]).unsqueeze(-1)
    b = pkg.Shrink(cq, "bayes")(rq, cvq)
    assert b.shape == rq.shape and b.dtype == torch.complex64
    assert np.abs(b.cpu().numpy().reshape(6, 16) - g["bayes"]).max() < 2e-6
    e, dxdr = pkg.Shrink(co, "shrinkOOK")(ro, cvo)
    assert e.dtype == torch.float32 and dxdr.dim() == 0
    assert np.abs(e.cpu().numpy().reshape(6, 16) - g["ook_exp"]).max() < 2e-6
    assert abs(float(dxdr) - float(g["ook_dxdr"])) < 2e-6 * abs(float(g["ook_dxdr"]))
    E, V = pkg.Shrink(co, "shrinkOOK").sw_shrinkOOK(ro, cvo)
    assert np.abs(E.cpu().numpy().reshape(6, 16) - g["sw_exp"]).max() < 2e-6
    assert np.abs(V.cpu().numpy().reshape(6, 16) - g["sw_var"]).max() < 2e-6
    # scalar cov, larger ragged size, against the oracle
    rng = np.random.default_rng(3)
    r = (rng.normal(size=(6, 1000 * 16)) + 1j * rng.normal(size=(6, 1000 * 16))).astype(np.complex64) * 0.7
    cbig = pkg.Config(16000, 2, 8, 1, 1, batch=6, generator_mode='random', alphabet='QPSK', channel_profile='uniform', device=DEV)
    got = pkg.Shrink(cbig, "bayes")(t(r).unsqueeze(-1), torch.tensor(0.3)).cpu().numpy().reshape(r.shape)
    assert np.abs(got - ao.shrink_bayes(r, np.float32(0.3), cbig.symbols, cbig.P0, cbig.Ps)).max() < 2e-6
    with pytest.raises(UnboundLocalError):
        pkg.Shrink(cq, "shrink")(rq, cvq)
    with pytest.raises(AttributeError):
        pkg.Shrink(cq, "lasso")(rq, cvq)


ISI_TAPS = ["bamp_isi_cyc", "bamp_isi_trunc", "bamp_isi_big"]


@pytest.mark.parametrize("name", ISI_TAPS)
@pytest.mark.parametrize("exp", ["f64", "f32"])
def test_bamp_structured_operator_matches_reference_goldens(name, exp):
    """ampsm_bamp_detect_taps (block-convolution operator from the Lh tap matrices; cyclic / truncated / tail layouts of
    channel.py:56-72, 89-91; one- and four-slot tiles) against the reference run on the dense matrix: exit iteration
    within 1, trajectories within the tolerances of parity_utils, every error count identical.  The dense kernel on the
    rebuilt matrix must take the same decisions, and BAMP.forward on a dense block-Toeplitz matrix must route itself
    through the structured operator."""
    from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps, taps_from_matrix
    g = load_golden(name)
    F, N = g["x"].shape
    cyclic = g["meta"]["kwargs"].get("trunc") == "cyclic" and g["meta"]["matrix"] == "channel"
    for snr_db in sorted(set(g["snr_db"].tolist())):
        sel = np.nonzero(g["snr_db"] == snr_db)[0]
        cfg = config_from_meta(g["meta"], batch=len(sel), device=DEV)
        idx = (g["idx"][sel].reshape(len(sel), -1) + (np.arange(len(sel)) * N)[:, None]).reshape(-1)
        args = (t(g["y"][sel]).unsqueeze(-1), 10 ** (snr_db / 10), t(g["x"][sel]).unsqueeze(-1), g["sym"][sel].reshape(-1), idx)
        amp = pkg.BAMP(cfg, trajectory=True, exp=exp, shift="reference" if exp == "f64" else "section")
        taps = t(g["H"][sel])                                            # (frames, Lh, Nr, Nt)
        d = amp.detect_taps(taps, *args, cyclic=cyclic)
        iters = d.iters.cpu().numpy()
        assert np.abs(iters - g["iters"][sel]).max() <= 1, (iters, g["iters"][sel])
        per_frame = np.abs(d.xmmse.cpu().numpy().reshape(len(sel), N) - g["xmmse"][sel]).max(axis=1)
        assert per_frame.max() < 1e-2 and np.median(per_frame) < 1e-4
        tr = d.traj.cpu().numpy()
        check_trajectory(name + ".tau", tr[:, :, 0], g["tau"][sel])
        check_trajectory(name + ".var", tr[:, :, 1], g["varm"][sel])
        xmap = d.xmap.cpu().numpy().reshape(len(sel), N)
        assert decision_mismatch_frames(cfg, xmap, g["xmap"][sel]).size == 0
        want = counters_for(cfg, g["xmap"][sel], g["xmmse"][sel], g["x"][sel], g["sym"][sel], idx)
        assert_counts_equal(f"{name}@{snr_db}dB taps", d.counters_dict(), want)
        # dense kernel on the rebuilt matrix, and the automatic routing of forward()
        H = matrix_from_taps(taps, cfg.Lin, cfg.Lout, cyclic)
        st = taps_from_matrix(H, cfg)
        assert st is not None and st[1] == cyclic and torch.equal(st[0], taps)
        dense = pkg.BAMP(cfg, exp=exp, shift="reference" if exp == "f64" else "section", structured=False).detect(H, *args)
        assert_counts_equal(f"{name}@{snr_db}dB dense", dense.counters_dict(), want)
        assert np.abs(dense.xmmse.cpu().numpy().reshape(len(sel), N) - g["xmmse"][sel]).max() < 1e-2
        auto = pkg.BAMP(cfg, exp=exp, shift="reference" if exp == "f64" else "section").detect(H, *args)
        assert torch.equal(auto.xmmse, d.xmmse) and torch.equal(auto.iters, d.iters)


def test_bamp_structured_operator_shared_taps_and_unstructured_matrix():
    """One tap set shared by the frames of a call (taps_frame_stride = 0) equals per-frame copies; a dense matrix that is
    not block-Toeplitz stays on the dense kernel."""
    from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps, taps_from_matrix
    g = load_golden("bamp_isi_cyc")
    F, N = g["x"].shape
    cfg = config_from_meta(g["meta"], batch=F, device=DEV)
    taps = t(g["H"][0])
    H = matrix_from_taps(taps, cfg.Lin, cfg.Lout, True)
    x = t(g["x"])
    y = (H @ x.T).T.contiguous() + 0.05 * t(g["y"])
    idx = global_idx(g, N)
    amp = pkg.BAMP(cfg)
    a = amp.detect_taps(taps, y.unsqueeze(-1), 10.0, x.unsqueeze(-1), g["sym"].reshape(-1), idx, cyclic=True)
    b = amp.detect_taps(taps.expand(F, *taps.shape).contiguous(), y.unsqueeze(-1), 10.0, x.unsqueeze(-1), g["sym"].reshape(-1),
                        idx, cyclic=True)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert torch.equal(a.xmmse, b.xmmse) and torch.equal(a.iters, b.iters)
    assert_counts_equal("shared vs per-frame taps", ca, cb)
    assert ca["sqerr"] == pytest.approx(cb["sqerr"], rel=1e-12)                  # double atomics: order of the CTAs only
    Hb = H.clone()
    Hb[0, -1] += 0.25
    assert taps_from_matrix(Hb, cfg) is None
    c = amp.detect(Hb, y.unsqueeze(-1), 10.0, x.unsqueeze(-1), g["sym"].reshape(-1), idx)
    assert c.counters_dict()["frames"] == F


def test_vamp_accepts_the_conj_view_factors_of_a_cuda_svd():
    """The reference's caller hands `torch.linalg.svd(A)` straight to VAMP (vamp_model.py:58-61).  On CUDA that Vh is a lazy
    conj-view (`is_conj()`), whose data_ptr() addresses un-conjugated memory: the host layer must materialise it.  C3
    shapes at 2 dB: the reference decodes every frame (40/40, probed on CPU); mis-read factors decode none."""
    F = 64
    cfg = pkg.Config(128, 4, 64, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                     device=DEV)
    np.random.seed(0)
    torch.manual_seed(0)
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    snr = 10 ** 0.2
    _, A = ch.generate_as_sparc()
    x, sym, idx = da.generate_message()
    y = A @ x + ch.awgn(snr)
    U, s, Vh = torch.linalg.svd(A, full_matrices=False)
    Vc = Vh.resolve_conj().contiguous()
    lazy = Vc.conj().mT.mH.conj().conj() if not Vh.is_conj() else Vh          # make sure a conj-view reaches the detector
    if not lazy.is_conj():
        lazy = Vc.conj().clone().conj()
    assert lazy.is_conj() and torch.equal(lazy.resolve_conj(), Vc)
    amp = pkg.VAMP(cfg)
    a = amp.detect(U, s, lazy, y, snr, x, sym, idx)
    b = amp.detect(U, s, Vc, y, snr, x, sym, idx)
    assert torch.equal(a.xmmse, b.xmmse) and torch.equal(a.iters, b.iters)
    assert a.counters_dict()["index_err"] == 0 and a.counters_dict()["nan_frames"] == 0
    # the same through forward() with a conj-view y and H for BAMP
    cfb = pkg.Config(64, 1, 32, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                     device=DEV)
    chb, dab = pkg.Channel(cfb), pkg.Data(cfb)
    H = chb.generate_channel()
    xb, sb, ib = dab.generate_message()
    yb = H @ xb + chb.awgn(10.0)
    Hl = H.conj().clone().conj()
    assert Hl.is_conj()
    c1 = pkg.BAMP(cfb).detect(H, yb, 10.0, xb, sb, ib)
    c2 = pkg.BAMP(cfb).detect(Hl, yb.conj().clone().conj(), 10.0, xb, sb, ib)
    assert torch.equal(c1.xmmse, c2.xmmse)
