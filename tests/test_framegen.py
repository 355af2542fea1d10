"""On-device frame generation (csrc/framegen.cuh): the Philox block function against the published
known-answer vectors, the generator kernel against its numpy restatement (oracle/framegen_oracle.py), chunk independence, and
the fused path (frames drawn inside the Jacobi SVD kernel, H never in HBM) against generate-then-detect, bit for bit."""
import numpy as np
import pytest
import torch

import amp_sparc_spatialmodulation_b200 as pkg
from oracle import framegen_oracle as fo

from parity_utils import INT_KEYS

DEV = "cuda:0"


def ints(c):
    return {k: c[k] for k in INT_KEYS}


def test_philox_block_function_matches_published_known_answers():
    """Random123's kat_vectors for philox4x32 with 10 rounds (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as
    1, 2, 3", SC'11): counter / key -> output."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        assert tuple(int(v) for v in fo.philox4x32_10(np.array(ctr), key)) == out


def test_oracle_frames_have_the_reference_distribution():
    """channel.py:55: entries CN(0, 1/Nr); data.py:84-85: one uniformly chosen antenna per section, uniformly chosen symbol;
    channel.py:113-115: noise CN(0, sigma^2)."""
    cfg = pkg.Config(16, 2, 8, 1, 1, batch=1, generator_mode='sparc', alphabet='QPSK', channel_profile='uniform', device='cpu')
    F, sigma2 = 600, 0.3
    H, y, x, lab, idx = fo.frames(7, 0, F, cfg.n, cfg.N, cfg.M, cfg.L, cfg.symbols, cfg.gray, 1 / cfg.Nr, sigma2)
    assert abs(np.mean(np.abs(H) ** 2) * cfg.Nr - 1) < 0.02 and abs(H.mean()) < 5e-3
    assert abs(np.mean(H.real ** 2) / np.mean(H.imag ** 2) - 1) < 0.03
    noise = y - np.einsum('fij,fj->fi', H, x)
    assert abs(np.mean(np.abs(noise) ** 2) / sigma2 - 1) < 0.05
    pos = idx.reshape(F, cfg.L) - (np.arange(F) * cfg.N)[:, None]
    assert ((pos // cfg.M) == np.arange(cfg.L)[None, :]).all() and (np.count_nonzero(x, axis=1) == cfg.L).all()
    ant = np.bincount((pos % cfg.M).ravel(), minlength=cfg.M) / pos.size
    assert np.abs(ant - 1 / cfg.M).max() < 0.04
    # two chunks of the stream are the stream
    H2 = fo.frames(7, 100, 5, cfg.n, cfg.N, cfg.M, cfg.L, cfg.symbols, cfg.gray, 1 / cfg.Nr, sigma2)[0]
    assert np.array_equal(H2, H[100:105])


@pytest.mark.gpu
@pytest.mark.parametrize("shape,alphabet,channel,method", [((64, 1, 32), '16QAM', 'iid', 'ar1'), ((64, 1, 32), 'QPSK', 'kronecker', 'ar1'),
                                                          ((64, 1, 32), 'QPSK', 'kronecker', 'roots'), ((16, 2, 8), 'QPSK', 'iid', 'ar1'),
                                                          ((32, 4, 24), 'QPSK', 'kronecker', 'roots'), ((32, 4, 24), 'QPSK', 'kronecker', 'ar1')])
def test_generator_kernel_matches_its_numpy_restatement(shape, alphabet, channel, method):
    Nt, Na, Nr = shape
    F, snr = 37, 10 ** 1.2
    cfg = pkg.Config(Nt, Na, Nr, 1, 1, batch=F, generator_mode='sparc', alphabet=alphabet, channel_profile='uniform', device=DEV)
    st = pkg.FrameStream(cfg, seed=0x1234567890ABCDEF, channel=channel, rho_t=0.7, rho_r=0.9, method=method)
    H, y, x, lab, idx = st.frames(1000, F, snr, frame_base=5)
    sigma2 = (cfg.Na / cfg.Nr) / snr
    Rr = st.Rr.cpu().numpy() if st.Rr is not None else None
    Rt = st.Rt.cpu().numpy() if st.Rt is not None else None
    Ho, yo, xo, labo, idxo = fo.frames(st.seed, 1000, F, cfg.n, cfg.N, cfg.M, cfg.L, cfg.symbols, cfg.gray, 1 / cfg.Nr, sigma2, Rr, Rt,
                                       frame_base=5, rho_r=st.rho_r, rho_t=st.rho_t)
    # integers exactly; floats to the accuracy of the device's fast log / sin / cos (2^-21 absolute on O(1) numbers)
    assert np.array_equal(lab.cpu().numpy(), labo) and np.array_equal(idx.cpu().numpy(), idxo)
    assert np.array_equal(x.cpu().numpy(), xo)
    scale = np.sqrt(1 / cfg.Nr)
    assert np.abs(H.cpu().numpy() - Ho).max() < 2e-5 * scale * 6
    assert np.abs(y.cpu().numpy() - yo).max() < 1e-4
    # any chunking of the stream gives the same frames, bit for bit
    Ha, ya, xa, laba, idxa = st.frames(1000, 20, snr, frame_base=5)
    Hb, yb, xb, labb, idxb = st.frames(1020, F - 20, snr, frame_base=25)
    assert torch.equal(torch.cat([Ha, Hb]), H) and torch.equal(torch.cat([ya, yb]), y) and torch.equal(torch.cat([xa, xb]), x)
    assert torch.equal(torch.cat([laba, labb]), lab) and torch.equal(torch.cat([idxa, idxb]), idx)
    # another seed is another stream
    other = pkg.FrameStream(cfg, seed=1, channel=channel, rho_t=0.7, rho_r=0.9, method=method).frames(1000, F, snr)[0]
    assert not torch.equal(other, H)


@pytest.mark.gpu
@pytest.mark.parametrize("channel,alphabet,snr_db", [('iid', '16QAM', 12.0), ('kronecker', 'QPSK', 6.0)])
def test_vamp_on_frames_drawn_inside_the_svd_kernel_equals_generate_then_detect(channel, alphabet, snr_db):
    """ampsm_vamp_detect_generated (the channel matrix never in HBM) against FrameStream.frames + detect_from_channel on the same
    frame numbers: identical estimates, exit iterations and counters -- and the same ground truth."""
    F, first = 5000, 12345
    cfg = pkg.Config(64, 1, 32, 1, 1, batch=F, generator_mode='sparc', iterations=20, alphabet=alphabet, channel_profile='uniform', device=DEV)
    st = pkg.FrameStream(cfg, seed=99, channel=channel, rho_t=0.7, rho_r=0.7)
    snr = 10 ** (snr_db / 10)
    H, y, x, lab, idx = st.frames(first, F, snr)
    amp = pkg.VAMP(cfg, outputs=True)
    a = amp.detect_from_channel(H, y, snr, x, lab, idx)
    b, xb, labb, idxb = amp.detect_generated(st, first, F, snr, return_truth=True)
    assert torch.equal(xb, x) and torch.equal(labb, lab) and torch.equal(idxb, idx)
    ca, cb = a.counters_dict(), b.counters_dict()
    assert ints(ca) == ints(cb) and ca["frames"] == F and ca["nan_frames"] == 0
    assert abs(ca["sqerr"] - cb["sqerr"]) <= 1e-10 * ca["sqerr"]          # float64 atomics: the order of the additions is free
    assert torch.equal(a.iters, b.iters) and torch.equal(a.xmmse, b.xmmse) and torch.equal(a.xmap, b.xmap)
    assert 0 < ca["index_err"] < F // 2                                   # a working detector on a non-trivial point


@pytest.mark.gpu
def test_generated_channel_moments_and_kronecker_covariance():
    """20 k generated 32 x 64 channels: entry variance 1/Nr, no correlation for 'iid'; for 'kronecker' the receive-side
    covariance E[H H^H] / tr(Rt) = Rr and the transmit-side E[H^H H] / tr(Rr) = Rt (rho = 0.7 / 0.9) to sampling accuracy."""
    F = 20000
    cfg = pkg.Config(64, 1, 32, 1, 1, batch=F, generator_mode='sparc', alphabet='QPSK', channel_profile='uniform', device=DEV)
    H = pkg.FrameStream(cfg, seed=3).frames(0, F, 10.0)[0]
    assert abs(float((H.abs() ** 2).mean()) * cfg.Nr - 1) < 5e-3 and float(H.mean().abs()) < 1e-3
    C = torch.einsum('fij,fkj->ik', H, H.conj()) / (F * cfg.N) * cfg.Nr
    assert float((C - torch.eye(cfg.n, device=DEV)).abs().max()) < 0.03
    for method in ('ar1', 'roots'):
        check_kronecker_covariance(cfg, F, method)


def check_kronecker_covariance(cfg, F, method):
    st = pkg.FrameStream(cfg, seed=3, channel='kronecker', rho_t=0.7, rho_r=0.9, method=method)
    H = st.frames(0, F, 10.0)[0]
    i = torch.arange(cfg.n, device=DEV)
    Rr = 0.9 ** (i[:, None] - i[None, :]).abs().float()
    j = torch.arange(cfg.N, device=DEV)
    Rt = 0.7 ** (j[:, None] - j[None, :]).abs().float()
    Cr = torch.einsum('fij,fkj->ik', H, H.conj()).real / (F * cfg.N) * cfg.Nr
    Ct = torch.einsum('fji,fjk->ik', H.conj(), H).real / (F * cfg.n) * cfg.Nr
    assert float((Cr - Rr).abs().max()) < 0.03 and float((Ct - Rt).abs().max()) < 0.04


@pytest.mark.gpu
def test_monte_carlo_driver_with_the_kernel_generator():
    """simulate.MonteCarlo(generator='kernel'): VAMP draws its frames inside the SVD kernel, BAMP from the generator kernel; both
    see the same stream, so their frame counts match and the FER falls with the SNR."""
    cfg = pkg.Config(64, 1, 32, 1, 1, batch=4096, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='uniform', device=DEV)
    for alg in ('vamp', 'bamp'):
        mc = pkg.MonteCarlo(cfg, alg, frames_per_point=10000, chunk=4096, seed=5, generator='kernel')
        lo = mc.run_point(0.0, 0)
        hi = mc.run_point(8.0, 1)
        assert lo["frames"] == hi["frames"] == 10000 and lo["nan_frames"] == 0
        assert hi["frame_err"] < lo["frame_err"] and lo["frame_err"] > 0
        again = pkg.MonteCarlo(cfg, alg, frames_per_point=10000, chunk=2500, seed=5, generator='kernel').run_point(0.0, 0)
        assert {k: again[k] for k in ("frame_err", "index_err", "symbol_err", "iters")} == {k: lo[k] for k in ("frame_err", "index_err", "symbol_err", "iters")}
