"""GPU tests at sizes the oracle cannot reach: size-independent properties of the CUDA path, the host-buffer entry
points, edge cases (empty / ragged frame counts, shared matrices).  Run on the B200 box: pytest -m gpu."""
import ctypes

import numpy as np
import pytest
import torch

import amp_sparc_spatialmodulation_b200 as pkg
from amp_sparc_spatialmodulation_b200 import _cabi
from parity_utils import INT_KEYS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def make_frames(cfg, frames, snr_db, seed, shared_H=False):
    """Synthetic frames on the device: i.i.d. CN(0, 1/Nr) channel, one active antenna per section, AWGN."""
    g = torch.Generator(device=DEV).manual_seed(seed)
    n, N, M, L = cfg.n, cfg.N, cfg.M, cfg.L
    shape = (n, N) if shared_H else (frames, n, N)
    H = torch.view_as_complex(torch.randn(*shape, 2, device=DEV, generator=g) * float(np.sqrt(1 / cfg.Nr / 2)))
    ant = torch.randint(0, M, (frames, L), device=DEV, generator=g)
    k = torch.randint(0, cfg.K, (frames, L), device=DEV, generator=g)
    sym = torch.as_tensor(cfg.symbols).to(DEV, torch.complex64)
    gray = torch.as_tensor(np.asarray(cfg.gray)).to(DEV, torch.int64)
    pos = ant + torch.arange(L, device=DEV) * M
    x = torch.zeros(frames, N, dtype=torch.complex64, device=DEV)
    x.scatter_(1, pos, sym[k])
    sigma2 = (cfg.Na / cfg.Nr) / 10 ** (snr_db / 10)
    noise = torch.view_as_complex(torch.randn(frames, n, 2, device=DEV, generator=g) * float(np.sqrt(sigma2 / 2)))
    y = (torch.einsum('ij,fj->fi', H, x) if shared_H else torch.einsum('fij,fj->fi', H, x)) + noise
    idx = (pos + torch.arange(frames, device=DEV)[:, None] * N).reshape(-1)
    return H, y.contiguous(), x, gray[k].reshape(-1).contiguous(), idx.contiguous()


def c2(frames, alphabet='16QAM', Na=1):
    return pkg.Config(64, Na, 32, 1, 1, batch=frames, generator_mode='sparc', iterations=20, alphabet=alphabet,
                      channel_profile='uniform', device=DEV)


def ints(c):
    return {k: c[k] for k in INT_KEYS}


@pytest.mark.parametrize("fast", ["fast"])
def test_fast_and_generic_kernels_agree_on_every_count_c2(fast):
    """Two independent kernels (register-resident FFMA2 + separable denoiser vs shared-memory float64-exponent one)
    must take identical hard decisions on 40k frames per SNR point."""
    F = 40000
    cfg = c2(F)
    for snr_db in (5.0, 15.0):
        H, y, x, lab, idx = make_frames(cfg, F, snr_db, seed=11)
        a = pkg.BAMP(cfg, kernel=fast, outputs=True).detect(H, y, 10 ** (snr_db / 10), x, lab, idx)
        b = pkg.BAMP(cfg, kernel='generic', exp='f64', outputs=True).detect(H, y, 10 ** (snr_db / 10), x, lab, idx)
        ca, cb = a.counters_dict(), b.counters_dict()
        ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
        # a frame that does not converge wanders chaotically: its exit iteration is not comparable between two
        # float32 evaluation orders (the reference behaves the same between two BLAS builds); all others agree
        assert (ia == ib).mean() > 0.995 and (np.abs(ia - ib) <= 1).mean() > 0.998
        # list (and bound) decision differences instead of hiding them: near-ties are the only legitimate cause
        diff = {k: (ca[k], cb[k]) for k in INT_KEYS if ca[k] != cb[k]}
        # (at 5 dB a few frames in 10^4 do not converge within 20 iterations; their final estimate is chaotic in float32)
        assert all(abs(u - v) <= max(3, 5e-4 * F) for u, v in diff.values()), diff
        assert ca["sqerr"] == pytest.approx(cb["sqerr"], rel=1e-3)
        d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
        assert float(d.median()) < 1e-5 and float(torch.quantile(d, 0.999)) < 5e-3     # all but the chaotic frames


def test_bamp_properties_at_scale():
    F = 300_001                                        # ragged on purpose: not a multiple of warps, CTAs or chunks
    cfg = c2(F)
    H, y, x, lab, idx = make_frames(cfg, F, 15.0, seed=5)
    amp = pkg.BAMP(cfg, outputs=False)
    c1 = amp.detect(H, y, 10 ** 1.5, x, lab, idx).counters_dict()
    c2_ = amp.detect(H, y, 10 ** 1.5, x, lab, idx).counters_dict()
    assert ints(c1) == ints(c2_) and c1["iters"] == c2_["iters"]                 # deterministic counts
    assert c1["frames"] == F and F <= c1["iters"] <= 20 * F and c1["nan_frames"] == 0
    assert c1["slot_err"] == c1["frame_err"] == c1["slot_err_first"] == c1["slot_err_last"]   # Lin = 1
    assert c1["index_err"] <= c1["frame_err"] <= c1["index_err"] + c1["symbol_err"]
    fixed = pkg.BAMP(cfg, outputs=False, early_exit=False).detect(H, y, 10 ** 1.5, x, lab, idx).counters_dict()
    assert fixed["iters"] == 20 * F
    # frame order cannot matter: reverse the frames (flat indices follow their frame)
    perm = torch.arange(F - 1, -1, -1, device=DEV)
    idx_p = idx[perm] - perm * cfg.N + torch.arange(F, device=DEV) * cfg.N
    cp = amp.detect(H[perm], y[perm], 10 ** 1.5, x[perm], lab[perm], idx_p).counters_dict()
    for k in INT_KEYS:
        if k != "index_bit_err":                        # the XOR of flat indices depends on the frame position (loss.py:168)
            assert cp[k] == c1[k], k
    # sharding by frame_base: two half calls add up to the whole call
    h = F // 2
    d1 = amp.detect(H[:h], y[:h], 10 ** 1.5, x[:h], lab[:h], idx[:h], frame_base=0)
    d2 = amp.detect(H[h:], y[h:], 10 ** 1.5, x[h:], lab[h:], idx[h:], frame_base=h)
    # index_bits_kept follows the frames of the CALL (loss.py:20), so compare everything but the truncated XOR count
    s = {k: d1.counters_dict()[k] + d2.counters_dict()[k] for k in INT_KEYS}
    for k in INT_KEYS:
        if k != "index_bit_err":
            assert s[k] == c1[k], k


def test_high_snr_qpsk_decodes_every_frame():
    F = 20000
    cfg = c2(F, alphabet='QPSK')
    H, y, x, lab, idx = make_frames(cfg, F, 30.0, seed=3)
    for kernel in ('fast', 'generic'):
        c = pkg.BAMP(cfg, kernel=kernel, outputs=False).detect(H, y, 10 ** 3.0, x, lab, idx).counters_dict()
        assert c["frame_err"] == 0 and c["index_bit_err"] == 0 and c["symbol_bit_err"] == 0 and c["nan_frames"] == 0


@pytest.mark.parametrize("fast,Na,alphabet", [("fast", 4, "QPSK"), ("fast", 2, "QPSK"),
                                              ("fast", 4, "16QAM")])
def test_multi_section_fast_shape_matches_generic(fast, Na, alphabet):
    """64 x 32, QPSK, Na = 4 / 2 (sections of 16 / 32 antennas: sub-warp and whole-warp section reductions)."""
    F = 20000
    cfg = c2(F, alphabet=alphabet, Na=Na)
    snr_db = 6.0 if alphabet == 'QPSK' else 18.0          # a regime where the frames converge (else the counts are chaotic)
    H, y, x, lab, idx = make_frames(cfg, F, snr_db, seed=8)
    a = pkg.BAMP(cfg, kernel=fast).detect(H, y, 10 ** (snr_db / 10), x, lab, idx).counters_dict()
    b = pkg.BAMP(cfg, kernel='generic', exp='f64').detect(H, y, 10 ** (snr_db / 10), x, lab, idx).counters_dict()
    diff = {k: (a[k], b[k]) for k in INT_KEYS if a[k] != b[k]}
    if alphabet == 'QPSK':
        # at most two frames may decide differently (near-ties in float32); a frame carries up to 4 sections of bit errors
        assert all(abs(u - v) <= (8 if k.endswith("bit_err") else 2) for k, (u, v) in diff.items()), diff
    else:
        # 16-QAM with several sections never converges for ~15 % of the sections at any SNR (the reference's decision metric
        # lacks the |s|^2 term): those frames wander chaotically in float32, so -- as for VAMP -- the
        # acceptance is two-sided: as close to the float64-exponent kernel as the generic kernel's own float32-exp mode
        c = pkg.BAMP(cfg, kernel='generic', exp='f32').detect(H, y, 10 ** (snr_db / 10), x, lab, idx).counters_dict()
        for k in INT_KEYS:
            slack = 2e-3 * F * cfg.L * (4 if k.endswith("bit_err") else 1)
            assert abs(a[k] - b[k]) <= abs(c[k] - b[k]) + slack, (k, a[k], b[k], c[k])


def test_shared_matrix_and_edge_frame_counts():
    cfg = c2(7)
    H, y, x, lab, idx = make_frames(cfg, 7, 12.0, seed=2, shared_H=True)
    for kernel in ('fast', 'generic'):
        d = pkg.BAMP(pkg.Config(64, 1, 32, 1, 1, batch=7, generator_mode='sparc', alphabet='16QAM', channel_profile='uniform',
                                device=DEV), kernel=kernel).detect(H, y, 10 ** 1.2, x, lab, idx)
        assert d.counters_dict()["frames"] == 7
        # one frame at a time through the same shared matrix gives the same estimates
        one = pkg.BAMP(c2(1), kernel=kernel).detect(H, y[3:4], 10 ** 1.2, x[3:4], lab[3:4], idx[3:4] - 3 * cfg.N)
        assert torch.equal(one.xmmse[0], d.xmmse[3]) and int(one.iters[0]) == int(d.iters[3])
    # zero frames: a no-op that leaves the counters untouched
    lib = _cabi.lib()
    counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=DEV)
    rc = lib.ampsm_bamp_detect(_cabi.make_problem(cfg, 1), _cabi.make_alphabet(cfg), 0, H.data_ptr(), 0, y.data_ptr(), 0.1, None,
                               None, None, None, None, None, None, None, None, counters.data_ptr(), None)
    assert rc == 0 and int(counters.abs().sum()) == 0
    # malformed problem: Na does not divide Nt -> AMPSM_EINVAL with a message, nothing launched
    bad = _cabi.make_problem(cfg, 1)
    bad.Na = 3
    rc = lib.ampsm_bamp_detect(bad, _cabi.make_alphabet(cfg), 1, H.data_ptr(), 0, y.data_ptr(), 0.1, None, None, None, None,
                               None, None, None, None, None, counters.data_ptr(), None)
    assert rc == -1 and b"Na" in lib.ampsm_last_error()


def test_host_entry_point_equals_device_entry_point():
    """ampsm_bamp_detect_host (chunked, double-buffered copies) against the device call on the same 30k frames."""
    F = 30000                                           # 17 KB per frame -> several 96 MiB chunks
    cfg = c2(F)
    H, y, x, lab, idx = make_frames(cfg, F, 10.0, seed=21)
    dev = pkg.BAMP(cfg, outputs=True).detect(H, y, 10.0, x, lab, idx)
    lib = _cabi.lib()
    hH, hy, hx, hl, hi = (t.cpu().contiguous() for t in (H, y, x, lab, idx))
    counters = np.zeros(_cabi.NUM_COUNTERS, dtype=np.int64)
    iters = np.zeros(F, dtype=np.int32)
    xmmse = np.zeros((F, cfg.N), dtype=np.complex64)
    rc = lib.ampsm_bamp_detect_host(_cabi.make_problem(cfg, F), _cabi.make_alphabet(cfg), F, hH.data_ptr(), cfg.n * cfg.N,
                                    hy.data_ptr(), float((cfg.Na / cfg.Nr) / 10.0), None, hx.data_ptr(), hl.data_ptr(), hi.data_ptr(),
                                    None, xmmse.ctypes.data, None, iters.ctypes.data, None, counters.ctypes.data, 0)
    _cabi.check(rc, "ampsm_bamp_detect_host")
    got, want = _cabi.counters_to_dict(counters), dev.counters_dict()
    for k in INT_KEYS + ["iters"]:
        assert got[k] == want[k], k
    assert np.array_equal(iters, dev.iters.cpu().numpy())
    assert np.array_equal(xmmse, dev.xmmse.cpu().numpy().reshape(F, cfg.N))


def test_vamp_and_scamp_host_entry_points():
    lib = _cabi.lib()
    # VAMP, shared factors, 64 frames
    cfg = pkg.Config(16, 2, 8, 3, 2, batch=64, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                     channel_truncation='tail', device=DEV)
    np.random.seed(0)
    torch.manual_seed(0)
    ccpu = pkg.Config(16, 2, 8, 3, 2, batch=64, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                      channel_truncation='tail', device='cpu')
    W, A = pkg.Channel(ccpu).generate_as_sparc()
    x, lab, idx = pkg.Data(ccpu).generate_message()
    snr = 10.0
    y = A @ x + pkg.Channel(ccpu).awgn(snr)
    U, s, Vh = (t.contiguous() for t in torch.linalg.svd(A, full_matrices=False))   # keep the contiguous copies alive
    y, x, W, A = y.contiguous(), x.contiguous(), W.contiguous(), A.contiguous()
    lab, idx = np.ascontiguousarray(lab), np.ascontiguousarray(idx)
    dev = pkg.VAMP(cfg).detect(U, s, Vh, y, snr, x, lab, idx)
    counters = np.zeros(_cabi.NUM_COUNTERS, dtype=np.int64)
    xm = np.zeros((64, cfg.N), np.complex64)
    prob = _cabi.make_problem(cfg, 64, R=Vh.shape[0])
    rc = lib.ampsm_vamp_detect_host(prob, _cabi.make_alphabet(cfg), 64, 0, U.contiguous().data_ptr(), 0, s.contiguous().data_ptr(), 0,
                                    Vh.contiguous().data_ptr(), 0, y.contiguous().data_ptr(), float((cfg.Na / cfg.Nr) / snr), None,
                                    float(cfg.Na / cfg.Nt), x.contiguous().data_ptr(), np.ascontiguousarray(lab).ctypes.data,
                                    np.ascontiguousarray(idx).ctypes.data, None, xm.ctypes.data, None, None, None,
                                    counters.ctypes.data, 0)
    _cabi.check(rc, "ampsm_vamp_detect_host")
    got, want = _cabi.counters_to_dict(counters), dev.counters_dict()
    assert all(got[k] == want[k] for k in INT_KEYS + ["iters", "nan_frames"])        # float64 sums: atomic order differs
    assert got["sqerr"] == pytest.approx(want["sqerr"], rel=1e-9)
    assert np.array_equal(xm, dev.xmmse.cpu().numpy().reshape(64, cfg.N))
    # SCAMP, same inputs (the host entry point takes the dense matrix: compare with the dense device path)
    dev = pkg.SCAMP(cfg, structured=False).detect(W, A, y, snr, x, lab, idx)
    counters[:] = 0
    rc = lib.ampsm_scamp_detect_host(_cabi.make_problem(cfg, 64), _cabi.make_alphabet(cfg), 64, W.contiguous().data_ptr(),
                                     A.contiguous().data_ptr(), y.contiguous().data_ptr(), float((cfg.Na / cfg.Nr) / snr), None,
                                     x.contiguous().data_ptr(), np.ascontiguousarray(lab).ctypes.data,
                                     np.ascontiguousarray(idx).ctypes.data, None, xm.ctypes.data, None, None, None,
                                     counters.ctypes.data, 0)
    _cabi.check(rc, "ampsm_scamp_detect_host")
    got, want = _cabi.counters_to_dict(counters), dev.counters_dict()
    assert all(got[k] == want[k] for k in INT_KEYS + ["iters", "nan_frames"])
    assert got["sqerr"] == pytest.approx(want["sqerr"], rel=1e-9)
    assert np.array_equal(xm, dev.xmmse.cpu().numpy().reshape(64, cfg.N))


def test_scamp_zero_tile_skipping_is_exact():
    """A coupled (band) design matrix: skipping its all-zero tiles must not change a single bit."""
    cfg = pkg.Config(64, 2, 16, 8, 3, batch=96, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform',
                     channel_truncation='tail', device='cpu')
    np.random.seed(4)
    torch.manual_seed(4)
    ch = pkg.Channel(cfg)
    W, A = ch.generate_as_sparc()
    x, lab, idx = pkg.Data(cfg).generate_message()
    snr = 10 ** 0.8
    y = A @ x + ch.awgn(snr)
    assert float((A == 0).float().mean()) > 0.5            # the band structure is there
    gcfg = pkg.Config(64, 2, 16, 8, 3, batch=96, generator_mode='sparc', iterations=20, alphabet='QPSK',
                      channel_profile='uniform', channel_truncation='tail', device=DEV)
    band = pkg.SCAMP(gcfg, structured=False).detect(W, A, y, snr, x, lab, idx)
    dense = pkg.SCAMP(gcfg, structured=False).detect(W, A + 0.0 * 1e-30, y, snr, x, lab, idx)    # same values
    tiny = A.clone()
    tiny[A == 0] = 1e-38 + 0j                               # no zero tile left, numerically the same matrix
    full = pkg.SCAMP(gcfg, structured=False).detect(W, tiny, y, snr, x, lab, idx)
    assert torch.equal(band.xmmse, dense.xmmse)
    assert torch.allclose(band.xmmse, full.xmmse, atol=1e-6) and torch.equal(band.iters, full.iters)
    cb, cf = band.counters_dict(), full.counters_dict()
    assert all(cb[k] == cf[k] for k in INT_KEYS + ["iters", "nan_frames"])


def c3(frames, alphabet='QPSK', Na=4):
    return pkg.Config(128, Na, 64, 1, 1, batch=frames, generator_mode='sparc', iterations=20, alphabet=alphabet,
                      channel_profile='uniform', device=DEV)


def svd_factors(H):
    """Thin SVD factors of a batch of wide matrices through the Hermitian eigenproblem of H H^H (float64), as the
    batched in-kernel Jacobi route does: H = U diag(s) Vh with s descending."""
    Us, ss, Vs = [], [], []
    for H1 in H.split(16384):                                # cuSOLVER's batched eigh rejects very large batches
        Hd = H1.to(torch.complex128)
        w, V = torch.linalg.eigh(Hd @ Hd.mH)                # ascending
        w, V = w.flip(-1), V.flip(-1)
        s = w.clamp_min(0).sqrt()
        Vh = (V.mH @ Hd) / s.unsqueeze(-1)
        Us.append(V.to(torch.complex64)), ss.append(s.to(torch.float32)), Vs.append(Vh.to(torch.complex64))
    return torch.cat(Us).contiguous(), torch.cat(ss).contiguous(), torch.cat(Vs).contiguous()


@pytest.mark.parametrize("alphabet,Na,snr_db", [("16QAM", 1, 12.0), ("QPSK", 1, 8.0), ("QPSK", 4, 6.0), ("QPSK", 2, 7.0)])
def test_vamp_fast_and_generic_kernels_agree(alphabet, Na, snr_db):
    """Register-resident VAMP kernel (one warp per frame, FFMA2, separable / table denoiser) against the generic
    shared-memory kernel with float64 exponents on 20k frames with per-frame SVD factors."""
    F = 20000
    cfg = c2(F, alphabet=alphabet, Na=Na)
    H, y, x, lab, idx = make_frames(cfg, F, snr_db, seed=21)
    U, s, Vh = svd_factors(H)
    snr = 10 ** (snr_db / 10)
    a = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    b = pkg.VAMP(cfg, kernel='generic', exp='f64', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    c = pkg.VAMP(cfg, kernel='generic', exp='f32', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    ca, cb, cc = a.counters_dict(), b.counters_dict(), c.counters_dict()
    assert ca["frames"] == cb["frames"] == F and ca["nan_frames"] == cb["nan_frames"] == 0
    ia, ib, ic = a.iters.cpu().numpy(), b.iters.cpu().numpy(), c.iters.cpu().numpy()
    # VAMP's sigma2_tilde is posterior tail mass: exit iterations of slowly converging frames are not comparable between
    # two float32 evaluation orders.  Two-sided acceptance: the register-resident kernel must be as
    # close to the float64-exponent kernel as the generic kernel's own float32-exp mode is.
    assert (ia == ib).mean() >= (ic == ib).mean() - 0.06, ((ia == ib).mean(), (ic == ib).mean())
    assert (np.abs(ia - ib) <= 1).mean() >= (np.abs(ic - ib) <= 1).mean() - 0.04
    assert abs(ia.mean() - ib.mean()) < 0.02 * ib.mean()
    # Net error counts: any two float32 evaluation orders of VAMP decide ~0.5 % of the frames differently, in both
    # directions (scripts/diag_vamp_flip.py on B200, 20k frames at 12 dB: fast vs f64-exp 103-117 frames, generic f32-exp
    # vs f64-exp 83-94, each split about evenly into better / worse), so the NET difference is a random walk of
    # ~sqrt(100) = 10 counts per sigma; 2e-3 F = 40 is four sigma.
    for k in INT_KEYS:
        slack = max(4, 2e-3 * F) * (4 if k.endswith("bit_err") else 1)
        assert abs(ca[k] - cb[k]) <= abs(cc[k] - cb[k]) + slack, (k, ca[k], cb[k], cc[k])
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    d32 = (c.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert float(d.median()) < 1e-5 and float(torch.quantile(d, 0.9)) <= max(1e-5, 10 * float(torch.quantile(d32, 0.9)))
    # deterministic, and the early-exit-disabled mode runs exactly T iterations
    a2 = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    assert torch.equal(a.xmmse, a2.xmmse) and ints(a2.counters_dict()) == ints(ca)
    fixed = pkg.VAMP(cfg, kernel='fast', outputs=False, early_exit=False).detect(U, s, Vh, y, snr, x, lab, idx).counters_dict()
    assert fixed["iters"] == 20 * F


@pytest.mark.parametrize("n,N", [(32, 64), (24, 64), (8, 16), (4, 8), (32, 32), (64, 128), (48, 64), (64, 64), (40, 128)])
def test_batched_jacobi_svd_reconstructs_and_matches_lapack(n, N):
    """ampsm_svd_batched: H = U diag(s) Vh to float32 accuracy, orthonormal factors, singular values equal to LAPACK's
    (torch.linalg.svdvals in float64), descending order; an ill-conditioned (Kronecker-correlated) batch included."""
    F = 3000 if n <= 32 else 600                       # more than 32 rows: the one-CTA-per-matrix kernel (64 x 128 = BASELINE config 3)
    g = torch.Generator(device=DEV).manual_seed(4)
    H = torch.view_as_complex(torch.randn(F, n, N, 2, device=DEV, generator=g) * float(np.sqrt(0.5 / n)))
    # second half: exponential correlation on both sides (rho = 0.9): condition numbers of 1e2 .. 1e3
    def corr_root(m, rho):
        R = rho ** (torch.arange(m, device=DEV)[:, None] - torch.arange(m, device=DEV)[None, :]).abs().double()
        w, V = torch.linalg.eigh(R)
        return (V * w.clamp_min(0).sqrt()) @ V.T
    Hc = (corr_root(n, 0.9).to(torch.complex128) @ H[F // 2:].to(torch.complex128) @ corr_root(N, 0.9).to(torch.complex128))
    H = torch.cat([H[:F // 2], Hc.to(torch.complex64)]).contiguous()
    U, s, Vh, sw = pkg.svd_batched(H, return_sweeps=True)
    rec = (U * s.unsqueeze(-2).to(torch.complex64)) @ Vh
    scale = H.abs().amax(dim=(1, 2))
    tol = 2e-5 if n <= 32 else 6e-5                    # float32 through ~200 (n <= 32) / ~600 (n = 64) rotations per row
    assert float(((rec - H).abs().amax(dim=(1, 2)) / scale).max()) < tol
    eye_n = torch.eye(n, dtype=torch.complex64, device=DEV)
    assert float((U.mH @ U - eye_n).abs().max()) < tol and float((Vh @ Vh.mH - eye_n).abs().max()) < tol
    ref = torch.linalg.svdvals(H.to(torch.complex128))
    assert float(((s.double() - ref).abs() / ref[:, :1]).max()) < 1.5 * tol
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    assert 2 <= int(sw.min()) and int(sw.max()) <= (14 if n <= 32 else 19)


def test_vamp_from_channel_equals_factor_entry_point():
    """detect_from_channel (device Jacobi SVD + iterations in one call) against factors from float64 eigh of H H^H fed
    to the factor entry point: same decisions on all but near-tie frames."""
    F = 20000
    cfg = c2(F)
    H, y, x, lab, idx = make_frames(cfg, F, 12.0, seed=31)
    snr = 10 ** 1.2
    U, s, Vh = svd_factors(H)
    a = pkg.VAMP(cfg, outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    b = pkg.VAMP(cfg, outputs=True).detect_from_channel(H, y, snr, x, lab, idx)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert cb["frames"] == F and cb["nan_frames"] == 0
    ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
    assert (ia == ib).mean() > 0.97
    for k in INT_KEYS:
        assert abs(ca[k] - cb[k]) <= max(4, 2e-3 * F) * (4 if k.endswith("bit_err") else 1), (k, ca[k], cb[k])
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert float(d.median()) < 1e-5


def test_vamp_from_channel_config3_equals_factor_entry_point():
    """BASELINE config 3 from the channel matrices: 64 x 128 Jacobi SVD (one CTA per matrix) + the four-warps-per-frame VAMP
    kernel in one call, against float64 eigh factors fed to the factor entry point."""
    F = 3000
    cfg = c3(F)
    H, y, x, lab, idx = make_frames(cfg, F, 4.0, seed=33)
    snr = 10 ** 0.4
    U, s, Vh = svd_factors(H)
    a = pkg.VAMP(cfg, outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    b = pkg.VAMP(cfg, outputs=True).detect_from_channel(H, y, snr, x, lab, idx)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert cb["frames"] == F and cb["nan_frames"] == 0
    ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
    assert (np.abs(ia - ib) <= 1).mean() > 0.95, (np.abs(ia - ib) <= 1).mean()
    for k in INT_KEYS:
        assert abs(ca[k] - cb[k]) <= max(6, 3e-3 * F * cfg.L) * (4 if k.endswith("bit_err") else 1), (k, ca[k], cb[k])
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert float(d.median()) < 1e-5


def test_monte_carlo_sweep_exports_reference_schema(tmp_path):
    """simulate.MonteCarlo (the GPU counterpart of Model.simulate, bamp_model.py:44-67): device-generated frames, one JSON
    per Eb/N0 point with the reference's keys, error rates falling with SNR, VAMP on Kronecker-correlated channels."""
    import json
    cfg = pkg.Config(64, 1, 32, 1, 1, batch=8192, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', device=str(DEV))
    mc = pkg.MonteCarlo(cfg, 'bamp', frames_per_point=20000, chunk=8192, path=str(tmp_path), device=DEV)
    pts = mc.simulate(start=-6.0, final=2.0, step=4.0, stop_fer=0.0)
    assert [p["EbN0dB"] for p in pts] == [-6.0, -2.0, 2.0] and all(p["frames"] == 20000 for p in pts)
    assert pts[0]["fer"] > pts[1]["fer"] > pts[2]["fer"] and pts[0]["fer"] > 0.1
    d = json.load(open(tmp_path / "-2.0.json"))
    for k in pkg.Loss.keys + ['T', 'EbN0dB', 'SNRdB', 'rate', 'C', 'ShannonLimitdB']:
        assert k in d, k
    assert d["fer"] == pytest.approx(pts[1]["fer"]) and d["T"] == pytest.approx(pts[1]["T"])
    # same seed, same counters; another chunking of the same pool changes the draws but not the statistics
    again = pkg.MonteCarlo(cfg, 'bamp', frames_per_point=20000, chunk=8192, device=DEV).run_point(-2.0, 1)
    assert again["frame_err"] == round(pts[1]["fer"] * 20000)
    vm = pkg.MonteCarlo(cfg, 'vamp', frames_per_point=12000, chunk=4096, channel='kronecker', rho_t=0.5, rho_r=0.5, device=DEV)
    c0, c1 = vm.run_point(-2.0, 0), vm.run_point(6.0, 1)
    assert c0["frames"] == c1["frames"] == 12000 and c0["nan_frames"] == 0
    assert c1["frame_err"] < c0["frame_err"]


@pytest.mark.parametrize("trunc", ["tail", "cyclic"])
def test_monte_carlo_isi_frames_run_through_the_structured_operator(trunc):
    """ISI frames (Lin = 8, Lh = 3) generated on the device as taps only and detected by ampsm_bamp_detect_taps: every frame
    counted, no NaN, error rates fall with SNR; the same frames through the dense kernel on the rebuilt block-Toeplitz
    matrices give identical counters."""
    from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps
    from amp_sparc_spatialmodulation_b200.simulate import device_frames
    cfg = pkg.Config(32, 2, 16, 8, 3, batch=2048, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='exponential', channel_truncation=trunc, device=str(DEV))
    mc = pkg.MonteCarlo(cfg, 'bamp', frames_per_point=5000, chunk=2048, device=DEV)
    c0, c1 = mc.run_point(-2.0, 0), mc.run_point(6.0, 1)
    assert c0["frames"] == c1["frames"] == 5000 and c0["nan_frames"] == c1["nan_frames"] == 0
    assert c1["index_err"] < c0["index_err"] and c0["index_err"] > 0
    gen = torch.Generator(device=DEV).manual_seed(11)
    taps, y, x, lab, idx = device_frames(cfg, 600, 2.0, gen)
    amp = pkg.BAMP(cfg, outputs=False)
    a = amp.detect_taps(taps, y, 2.0, x, lab, idx, cyclic=trunc == 'cyclic').counters_dict()
    H = matrix_from_taps(taps, cfg.Lin, cfg.Lout, trunc == 'cyclic')
    b = pkg.BAMP(cfg, outputs=False, structured=False).detect(H, y, 2.0, x, lab, idx).counters_dict()
    for k in INT_KEYS:
        assert abs(a[k] - b[k]) <= max(2, 0.002 * max(a[k], b[k])), (k, a[k], b[k])       # near-tie frames may flip
    assert a["iters"] == pytest.approx(b["iters"], rel=5e-3)


@pytest.mark.parametrize("shape", [(64, 2, 8, 8, 3), (128, 4, 16, 6, 2)])
def test_scamp_tensor_core_gemms_match_simt_path(shape, monkeypatch):
    """Batches of >= 128 frames run both SCAMP GEMMs on the tensor cores (tcgen05 kind::tf32, 3xTF32 split, scamp_tc.cu) and
    the fused one-warp-per-section denoiser; the SIMT float32 tiles with the generic denoiser (AMPSM_SCAMP_SIMT=1,
    AMPSM_SCAMP_GENERIC_DENOISER=1 -- the path the golden fixtures pin) are the comparison: same exits, same decisions, estimates to float32
    rounding.  Ragged on purpose: frames not a multiple of 128, outputs not a multiple of 64."""
    Nt, Na, Nr, Lin, Lh = shape
    F = 300
    cfg = pkg.Config(Nt, Na, Nr, Lin, Lh, batch=F, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', channel_truncation='tail', device=str(DEV))
    np.random.seed(3)
    torch.manual_seed(3)
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    W, A = ch.generate_as_sparc()
    x, sym, idx = da.generate_message()
    snr = 10 ** ((5.0 + 10 * np.log10(cfg.code_rate)) / 10)
    y = A @ x + ch.awgn(snr)
    monkeypatch.delenv("AMPSM_SCAMP_SIMT", raising=False)
    a = pkg.SCAMP(cfg, outputs=True, structured=False).detect(W, A, y, snr, x, sym, idx)
    ca = a.counters_dict()
    monkeypatch.setenv("AMPSM_SCAMP_SIMT", "1")
    monkeypatch.setenv("AMPSM_SCAMP_GENERIC_DENOISER", "1")      # the reference path: SIMT GEMM tiles + the generic denoiser
    b = pkg.SCAMP(cfg, outputs=True, structured=False).detect(W, A, y, snr, x, sym, idx)
    cb = b.counters_dict()
    monkeypatch.delenv("AMPSM_SCAMP_GENERIC_DENOISER")
    assert ca["frames"] == cb["frames"] == F and ca["nan_frames"] == 0
    ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
    assert (np.abs(ia - ib) <= 1).mean() > 0.99 and (ia == ib).mean() > 0.75      # psi = 1 - (~1): the exit test sits at float32 resolution
    # frames that met the exit test at the same iteration must agree to rounding; frames that run out of iterations without
    # converging amplify any rounding difference chaotically and are only counted
    conv = torch.as_tensor((ia == ib) & (ia < 20), device=DEV)
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert int(conv.sum()) > F // 2
    assert float(torch.quantile(d[conv], 0.98)) < 2e-4 and float(d[conv].median()) < 1e-5
    for k in INT_KEYS:
        assert abs(ca[k] - cb[k]) <= max(3, 0.02 * cb[k]) * (4 if k.endswith("bit_err") else 1), (k, ca[k], cb[k])


@pytest.mark.parametrize("alphabet,Na,snr_db", [("QPSK", 4, 2.0), ("QPSK", 8, 4.0), ("16QAM", 4, 10.0)])
def test_vamp_quad_and_generic_kernels_agree(alphabet, Na, snr_db):
    """Four-warps-per-frame register-resident VAMP kernel (csrc/vamp_quad.cu, the 128 x 64 shapes of BASELINE config 3)
    against the generic shared-memory kernel, per-frame SVD factors; same two-sided acceptance as the 64 x 32 kernels."""
    F = 6000
    cfg = c3(F, alphabet=alphabet, Na=Na)
    H, y, x, lab, idx = make_frames(cfg, F, snr_db, seed=31)
    U, s, Vh = svd_factors(H)
    snr = 10 ** (snr_db / 10)
    a = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    b = pkg.VAMP(cfg, kernel='generic', exp='f64', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    c = pkg.VAMP(cfg, kernel='generic', exp='f32', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    ca, cb, cc = a.counters_dict(), b.counters_dict(), c.counters_dict()
    assert ca["frames"] == cb["frames"] == F and ca["nan_frames"] == cb["nan_frames"] == 0
    ia, ib, ic = a.iters.cpu().numpy(), b.iters.cpu().numpy(), c.iters.cpu().numpy()
    assert (ia == ib).mean() >= (ic == ib).mean() - 0.06, ((ia == ib).mean(), (ic == ib).mean())
    assert (np.abs(ia - ib) <= 1).mean() >= (np.abs(ic - ib) <= 1).mean() - 0.04
    assert abs(ia.mean() - ib.mean()) < 0.02 * ib.mean()
    for k in INT_KEYS:
        slack = max(6, 2e-3 * F * cfg.L) * (4 if k.endswith("bit_err") else 1)      # see the 64 x 32 test for the 2e-3
        assert abs(ca[k] - cb[k]) <= abs(cc[k] - cb[k]) + slack, (k, ca[k], cb[k], cc[k])
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    d32 = (c.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert float(d.median()) < 1e-5 and float(torch.quantile(d, 0.9)) <= max(1e-5, 10 * float(torch.quantile(d32, 0.9)))
    # deterministic; the early-exit-disabled mode runs exactly T iterations; shared factors = per-frame factors
    a2 = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U, s, Vh, y, snr, x, lab, idx)
    assert torch.equal(a.xmmse, a2.xmmse) and ints(a2.counters_dict()) == ints(ca)
    fx = pkg.VAMP(cfg, kernel='fast', outputs=False, early_exit=False).detect(U, s, Vh, y, snr, x, lab, idx)
    assert fx.counters_dict()["iters"] == 20 * F
    sh = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U[0], s[0], Vh[0], y[:64], snr, x[:64], lab[:64], idx[:64])
    pf = pkg.VAMP(cfg, kernel='fast', outputs=True).detect(U[:1].expand(64, -1, -1).contiguous(), s[:1].expand(64, -1).contiguous(),
                                                           Vh[:1].expand(64, -1, -1).contiguous(), y[:64], snr, x[:64], lab[:64], idx[:64])
    assert torch.equal(sh.xmmse, pf.xmmse)


@pytest.mark.parametrize("alphabet", ["16QAM", "QPSK"])
def test_fast_loss_degenerate_frames_match_generic(alphabet):
    """The fused Loss of the register-resident kernels (fast_loss2: corner shortcut for product-grid alphabets, tournament
    for everything else, REDUX arg-max on an order-preserving key) on inputs that defeat the shortcut: all-zero channels
    (xmap == 0 exactly: every metric is +-0, np.argmax takes flat index 0), NaN and Inf observations (first NaN wins),
    frames whose xmap has an exactly-zero real or imaginary part, mixed with ordinary frames.  Every integer count and the
    per-frame decisions must equal the generic kernel's (block_loss, the np.argmax restatement pinned by the Loss goldens)."""
    F = 4096
    cfg = c2(F, alphabet=alphabet)
    H, y, x, lab, idx = make_frames(cfg, F, 12.0, seed=77)
    H = H.clone()
    y = y.clone()
    H[0:64] = 0                                       # xmap = 0 exactly
    y[64:96] = float('nan')                           # NaN everywhere
    y[96:128, 0] = float('inf')                       # Inf -> NaN after the first products
    H[128:192] = H[128:192].real.to(torch.complex64)  # real channel ...
    y[128:192] = y[128:192].real.to(torch.complex64)  # ... and real observation: xmap.imag == 0 exactly
    snr = 10 ** 1.2
    res = {}
    for kernel in ('fast', 'generic'):
        d = pkg.BAMP(cfg, kernel=kernel, outputs=True).detect(H, y, snr, x, lab, idx)
        res[kernel] = (d.counters_dict(), d.xmap.reshape(F, -1).clone(), d.iters.clone())
    cf, cg = res['fast'][0], res['generic'][0]
    assert cf["nan_frames"] == cg["nan_frames"] >= 64
    # degenerate frames: identical counts frame group by frame group (run the groups alone)
    for lo, hi in ((0, 64), (64, 128), (128, 192)):
        sl = slice(lo, hi)
        n = hi - lo
        sub = c2(n, alphabet=alphabet)
        idx_s = idx[sl] - lo * cfg.N
        a = pkg.BAMP(sub, kernel='fast', outputs=False).detect(H[sl], y[sl], snr, x[sl], lab[sl], idx_s).counters_dict()
        b = pkg.BAMP(sub, kernel='generic', outputs=False).detect(H[sl], y[sl], snr, x[sl], lab[sl], idx_s).counters_dict()
        if lo < 128:
            assert ints(a) == ints(b), (lo, hi, ints(a), ints(b))
        else:
            # real channel and observation: the imaginary part of xmap is pure rounding (or exactly zero), so the metrics of
            # conjugate symbols nearly tie and two float32 evaluation orders may pick different ones in a few frames
            assert all(abs(a[k] - b[k]) <= 3 * (4 if k.endswith("bit_err") else 1) for k in INT_KEYS), (ints(a), ints(b))
    # the whole call: ordinary frames may differ by float32 near-ties only
    diff = {k: (cf[k], cg[k]) for k in INT_KEYS if cf[k] != cg[k]}
    assert all(abs(u - v) <= 3 * (4 if k.endswith("bit_err") else 1) for k, (u, v) in diff.items()), diff


def test_fast_kernel_falls_back_on_misaligned_loss_inputs():
    """The register-resident kernels stage x_true with 16-byte cp.async; an x_true that is only 8-byte aligned must take the
    generic kernel ('auto') with the same result, and be refused by kernel='fast'."""
    F = 2048
    cfg = c2(F)
    H, y, x, lab, idx = make_frames(cfg, F, 15.0, seed=9)
    buf = torch.zeros(F * cfg.N + 1, dtype=torch.complex64, device=DEV)
    xm = buf[1:].view(F, cfg.N)                       # data_ptr() % 16 == 8
    xm.copy_(x)
    assert xm.data_ptr() % 16 == 8
    ref = pkg.BAMP(cfg, outputs=False).detect(H, y, 10 ** 1.5, x, lab, idx).counters_dict()
    got = pkg.BAMP(cfg, outputs=False).detect(H, y, 10 ** 1.5, xm, lab, idx).counters_dict()
    diff = {k: (got[k], ref[k]) for k in INT_KEYS if got[k] != ref[k]}
    assert all(abs(u - v) <= 2 for u, v in diff.values()), diff     # generic vs fast kernel: near-ties only
    with pytest.raises(Exception):
        pkg.BAMP(cfg, kernel='fast', outputs=False).detect(H, y, 10 ** 1.5, xm, lab, idx)


@pytest.mark.parametrize("shape,F,trunc,alphabet", [((64, 2, 8, 8, 3), 300, 'tail', 'QPSK'), ((128, 4, 16, 6, 2), 150, 'tail', 'QPSK'),
                                                     ((64, 2, 8, 8, 3), 130, 'trunc', 'QPSK'), ((48, 2, 6, 5, 2), 140, 'tail', 'QPSK'),
                                                     ((128, 8, 32, 16, 3), 37, 'tail', 'QPSK'), ((64, 2, 8, 8, 3), 160, 'tail', 'QPSK-generic'), ((64, 2, 8, 8, 3), 160, 'tail', 'QPSK-exact'),
                                                     ((128, 2, 16, 6, 2), 140, 'tail', 'BPSK')])
def test_scamp_structured_path_matches_dense_path(shape, F, trunc, alphabet, monkeypatch):
    """The structured tensor-core kernels (design matrix applied from its taps: tensor TMA, tcgen05, csrc/scamp_st.cu) against
    the dense tensor-core / SIMT kernels reading the full matrix, same frames: estimates of frames that exit at the same
    iteration agree to float32 rounding, decisions agree on all but near-tie / non-converged frames.  Shapes cover ragged tiles
    (Lin = 6, 5: 126 / 125 rows per tile), a truncated channel (zero padding blocks of Zs), Nt and Lh Nr that are not multiples of
    the MMA tile (48 columns, 12 / 16 / 24 reduction rows), a batch smaller than one tile, and alphabets that take the generic
    (compensated float32 exponent) branch of the fused denoiser instead of the QPSK factorisation."""
    Nt, Na, Nr, Lin, Lh = shape
    if alphabet == 'QPSK-generic':        # the compensated float32 exponent branch (what an alphabet outside {0, +-1, +-j} takes)
        monkeypatch.setenv("AMPSM_SCAMP_NO_EXACT", "1")
        alphabet = 'QPSK'
    if alphabet == 'QPSK-exact':          # the exact-product branch without the axis-symbol shortcut the reference's QPSK table takes
        monkeypatch.setenv("AMPSM_SCAMP_NO_AXIS4", "1")
        alphabet = 'QPSK'
    cfg = pkg.Config(Nt, Na, Nr, Lin, Lh, batch=F, generator_mode='sparc', iterations=20, alphabet=alphabet,
                     channel_profile='uniform', channel_truncation=trunc, device=str(DEV))
    np.random.seed(13)
    torch.manual_seed(13)
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    W, A = ch.generate_as_sparc()
    x, sym, idx = da.generate_message()
    ebn0 = 6.0
    snr = 10 ** ((ebn0 + 10 * np.log10(cfg.code_rate)) / 10)
    y = A @ x + ch.awgn(snr)
    st = pkg.SCAMP(cfg, outputs=True)
    assert st._taps_of(A.to(DEV) if not A.is_cuda else A) is not None
    a = st.detect(W, A, y, snr, x, sym, idx)
    b = pkg.SCAMP(cfg, outputs=True, structured=False).detect(W, A, y, snr, x, sym, idx)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert ca["frames"] == cb["frames"] == F and ca["nan_frames"] == cb["nan_frames"] == 0
    ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
    assert (np.abs(ia - ib) <= 1).mean() > 0.97 and (ia == ib).mean() > 0.75, ((np.abs(ia - ib) <= 1).mean(), (ia == ib).mean())
    # frames that met the exit test at the same iteration agree to float32 rounding and decide alike; frames that run out of
    # iterations wander chaotically in any two float32 evaluation orders and are only counted
    conv_np = (ia == ib) & (ia < 20)
    conv = torch.as_tensor(conv_np, device=DEV)
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert int(conv.sum()) >= 8, int(conv.sum())
    assert float(torch.quantile(d[conv], 0.98)) < 2e-4 and float(d[conv].median()) < 1e-5, (float(d[conv].median()), float(d[conv].max()))
    from parity_utils import decision_mismatch_frames
    bad = decision_mismatch_frames(cfg, a.xmap.cpu().numpy().reshape(F, -1), b.xmap.cpu().numpy().reshape(F, -1))
    assert np.intersect1d(bad, np.nonzero(conv_np)[0]).size <= max(1, 0.01 * conv_np.sum()), bad
    for k in INT_KEYS:
        assert abs(ca[k] - cb[k]) <= max(8, 0.1 * cb[k]) * (4 if k.endswith("bit_err") else 1), (k, ca[k], cb[k])


def test_vamp_complex128_register_kernel_matches_generic_double_kernel(monkeypatch):
    """complex128 factors at BASELINE config 3 (VAMP 128 x 64, Na = 4): the register-resident DFMA kernel (csrc/vamp_dbl.cu)
    against the generic float64 kernel (Vh in shared memory) on 1500 frames with per-frame factors, plus a shared-factor call:
    same exit iterations, estimates to 1e-6 (var is rounded to float32 every iteration, vamp.py:119), identical counters."""
    F = 1500
    cfg = c3(F)
    H, y, x, lab, idx = make_frames(cfg, F, 3.0, seed=41)
    U, s, Vh = torch.linalg.svd(H.to(torch.complex128), full_matrices=False)
    U, s, Vh, yd = U.contiguous(), s.contiguous(), Vh.contiguous(), y.to(torch.complex128)
    snr = 10 ** 0.3
    a = pkg.VAMP(cfg, outputs=True).detect(U, s, Vh, yd, snr, x, lab, idx)
    monkeypatch.setenv("AMPSM_VAMP_DBL_GENERIC", "1")
    b = pkg.VAMP(cfg, outputs=True).detect(U, s, Vh, yd, snr, x, lab, idx)
    monkeypatch.delenv("AMPSM_VAMP_DBL_GENERIC")
    ca, cb = a.counters_dict(), b.counters_dict()
    assert ca["frames"] == cb["frames"] == F and ca["nan_frames"] == cb["nan_frames"] == 0
    ia, ib = a.iters.cpu().numpy(), b.iters.cpu().numpy()
    assert (ia == ib).mean() > 0.995, (ia == ib).mean()
    same = torch.as_tensor(ia == ib, device=DEV)
    d = (a.xmmse - b.xmmse).abs().reshape(F, -1).amax(dim=1)
    assert float(d[same].max()) < 1e-5 and float(d.median()) < 1e-7
    for k in INT_KEYS:
        assert abs(ca[k] - cb[k]) <= 2 * (4 if k.endswith("bit_err") else 1), (k, ca[k], cb[k])
    sh = pkg.VAMP(cfg, outputs=True).detect(U[0], s[0], Vh[0], yd[:70], snr, x[:70], lab[:70 * cfg.L], idx[:70 * cfg.L])
    pf = pkg.VAMP(cfg, outputs=True).detect(U[:1].expand(70, -1, -1).contiguous(), s[:1].expand(70, -1).contiguous(),
                                            Vh[:1].expand(70, -1, -1).contiguous(), yd[:70], snr, x[:70], lab[:70 * cfg.L], idx[:70 * cfg.L])
    assert torch.equal(sh.xmmse, pf.xmmse)


def test_full_size_snr_point_is_the_sum_of_its_frames():
    """BASELINE config 2 at its full size -- 1 048 576 frames in ONE call (17.9 GB of channel matrices drawn by the generator
    kernel) -- through size-independent properties: the call is deterministic, every frame's estimate, exit iteration and decision
    equal what the same frame gives in a 4 096-frame call of its own (frames are independent: nothing may depend on the launch
    size, the grid-stride assignment or the position in the batch), and the counters of the whole point equal the sum over 16
    chunk calls.  The oracle meets these kernels on 10 k-frame subsets in tests/test_gpu_oracle_scale.py."""
    F, snr = 1 << 20, 10 ** 1.5
    cfg = c2(F)
    st = pkg.FrameStream(cfg, seed=2024)
    H, y, x, lab, idx = st.frames(0, F, snr)
    amp = pkg.BAMP(cfg, outputs=True)
    whole = amp.detect(H, y, snr, x, lab, idx)
    cw = whole.counters_dict()
    assert cw["frames"] == F and cw["nan_frames"] == 0 and F <= cw["iters"] <= 20 * F
    again = amp.detect(H, y, snr, x, lab, idx)
    assert torch.equal(again.xmmse, whole.xmmse) and torch.equal(again.iters, whole.iters) and ints(again.counters_dict()) == ints(cw)
    for lo in (0, 123_456, F - 4096):                  # windows at both ends and at an odd offset
        sl = slice(lo, lo + 4096)
        small = amp.detect(H[sl], y[sl], snr, x[sl], lab[sl], idx[sl] - lo * cfg.N)
        assert torch.equal(small.xmmse, whole.xmmse[sl]) and torch.equal(small.xmap, whole.xmap[sl])
        assert torch.equal(small.iters, whole.iters[sl]) and torch.equal(small.var, whole.var[sl])
    tot = {k: 0 for k in INT_KEYS + ["iters"]}
    step = F // 16
    quiet = pkg.BAMP(cfg, outputs=False)
    for lo in range(0, F, step):
        sl = slice(lo, lo + step)
        c = quiet.detect(H[sl], y[sl], snr, x[sl], lab[sl], idx[sl], frame_base=lo).counters_dict()
        for k in tot:
            tot[k] += c[k]
    for k in tot:
        if k != "index_bit_err":                        # its truncation follows the frames of the CALL (loss.py:20)
            assert tot[k] == cw[k], k
    del H, whole, again
    # VAMP at the same size with the frames drawn inside the SVD kernel: the point is the sum of its chunks, and a window of it
    # equals the same frame numbers detected on their own
    vamp = pkg.VAMP(cfg, outputs=False)
    vw = vamp.detect_generated(st, 0, F, snr).counters_dict()
    assert vw["frames"] == F and vw["nan_frames"] == 0
    vt = {k: 0 for k in INT_KEYS + ["iters"]}
    for lo in range(0, F, step):
        c = vamp.detect_generated(st, lo, step, snr, frame_base=lo).counters_dict()
        for k in vt:
            vt[k] += c[k]
    for k in vt:
        if k != "index_bit_err":
            assert vt[k] == vw[k], k
    a = pkg.VAMP(cfg, outputs=True).detect_generated(st, 0, 1 << 17, snr)
    b = pkg.VAMP(cfg, outputs=True).detect_generated(st, 100_000, 2048, snr)
    assert torch.equal(b.xmmse, a.xmmse[100_000:102_048]) and torch.equal(b.iters, a.iters[100_000:102_048])
