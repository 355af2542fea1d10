"""Host-side mirror of the reference API: Config attributes, generator RNG order, Loss bookkeeping, C-ABI exports."""
import os
import re

import numpy as np
import pytest
import torch

import amp_sparc_spatialmodulation_b200 as pkg
from amp_sparc_spatialmodulation_b200 import _cabi
from conftest import ROOT, config_from_meta, load_golden


def test_config_derived_attributes():
    c = pkg.Config(64, 1, 32, 1, 1, batch=7, generator_mode='sparc', iterations=20, alphabet='16QAM',
                   channel_profile='uniform', device='cpu')
    assert (c.B, c.Nt, c.Na, c.Nr, c.Lin, c.Lh, c.Lout, c.K) == (7, 64, 1, 32, 1, 1, 1, 16)
    assert (c.M, c.L, c.n, c.N, c.Mc, c.Mr, c.Lc, c.Lr) == (64, 1, 32, 64, 64, 32, 1, 1)
    assert c.symbol_bits == 4 and c.index_bits == 6.0
    assert c.code_rate == pytest.approx(np.log2(64 * 16) / 32)
    assert c.Ps == pytest.approx(1 / 64 / 16) and c.P0 == pytest.approx(63 / 64)
    # the reference's 16-QAM table: -1+3j twice, 1-3j missing, normalised over the listed points (config.py:112-117)
    s = c.symbols * np.sqrt(np.mean(np.abs(c.symbols * np.sqrt(10)) ** 2) / 10) * np.sqrt(10)
    assert np.isclose(np.mean(np.abs(c.symbols) ** 2), 1.0)
    assert c.symbols[13] == c.symbols[14] and c.gray[13] == 3 and c.gray[14] == 6
    assert not np.any(np.isclose(s / abs(s[0].real), 1 - 3j))
    assert c.name == '16QAM,sparc/uniform,trunc/Nt=64,Na=1,Nr=32,Lh=1,Lin=1'
    t = pkg.Config(128, 8, 24, 20, 3, generator_mode='sparc', alphabet='OOK', channel_truncation='tail', device='cpu')
    assert t.Lout == 22 and t.n == 24 * 22 and t.ISI and t.K == 1 and t.symbol_bits == 0 and not t.modulated


def test_config_asserts_like_reference():
    with pytest.raises(AssertionError):
        pkg.Config(8, 1, 4, 1, 0)
    with pytest.raises(AssertionError):
        pkg.Config(8, 1, 4, 1, 1, alphabet='64QAM')
    with pytest.raises(AssertionError):
        pkg.Config(8, 3, 4, 1, 1, generator_mode='sparc')
    with pytest.raises(AssertionError):
        pkg.Config(8, 1, 4, 1, 1, channel_truncation='none')


@pytest.mark.parametrize("name", ["bamp_c1", "bamp_isi", "bamp_seg"])
def test_generators_reproduce_reference_draws(name):
    """Same seeds + same RNG call order => inputs bit-identical to the reference's Channel/Data (golden H, x, y)."""
    g = load_golden(name)
    meta = g["meta"]
    cfg = config_from_meta(meta)
    np.random.seed(meta["seed"])
    torch.manual_seed(meta["seed"])
    ch, da = pkg.Channel(cfg), pkg.Data(cfg)
    f = 0
    for snr_db in sorted(set(g["snr_db"].tolist())):
        snr = 10 ** (snr_db / 10)
        for _ in range(int((g["snr_db"] == snr_db).sum())):
            H = ch.generate_channel() if meta.get("matrix", "channel") == "channel" else ch.generate_as_sparc()[1]
            x, s, i = da.generate_message()
            y = H @ x + ch.awgn(snr)
            assert np.array_equal(H.numpy(), g["H"][f])
            assert np.array_equal(x.numpy().ravel(), g["x"][f])
            assert np.array_equal(s, g["sym"][f]) and np.array_equal(i, g["idx"][f])
            assert np.allclose(y.numpy().ravel(), g["y"][f], rtol=0, atol=1e-6)
            f += 1
    assert f == g["H"].shape[0]


def test_scamp_generator_reproduces_reference():
    g = load_golden("scamp_small")
    cfg = config_from_meta(g["meta"])
    np.random.seed(g["meta"]["seed"])
    torch.manual_seed(g["meta"]["seed"])
    W, A = pkg.Channel(cfg).generate_as_sparc()
    assert np.array_equal(W.numpy(), g["W"][0]) and np.array_equal(A.numpy(), g["A"][0])


def test_loss_rates_and_bookkeeping():
    g = load_golden("loss_qpsk")
    F = g["x"].shape[0]
    cfg = config_from_meta(g["meta"], batch=F)
    from oracle import loss_oracle as lo
    c = lo.error_counters(g["xmap"], g["xmmse"], g["x"], g["sym"], g["idx"], cfg.symbols, cfg.gray,
                          dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin))
    L = pkg.Loss(cfg)
    L.record(c, 5)
    for k, want in zip(L.keys, g["loss"]):
        assert float(L.loss[k]) == pytest.approx(want, rel=1e-5, abs=1e-12), k
    assert L.loss['T'] == 5
    acc = pkg.Loss(cfg)
    acc.accumulate(L)
    acc.accumulate(L)
    acc.average(2)
    assert float(acc.loss['fer']) == pytest.approx(float(L.loss['fer'])) and acc.loss['T'] == 5
    L.dump()
    assert L.loss == {}


def test_loss_export_schema(tmp_path):
    cfg = pkg.Config(8, 1, 4, 1, 1, batch=4, generator_mode='sparc', alphabet='QPSK', channel_profile='uniform', device='cpu')
    L = pkg.Loss(cfg)
    c = dict(frames=4, frame_err=1, slot_err=1, slot_err_first=1, slot_err_mid=1, slot_err_last=1, index_err=1, symbol_err=0,
             index_bit_err=2, symbol_bit_err=0, iters=20, nan_frames=0, sqerr=0.5, sqerr_first=0.5, sqerr_mid=0.5, sqerr_last=0.5)
    L.record(c, 5)
    L.export(3.0, 6.0, str(tmp_path))
    import json
    d = json.load(open(tmp_path / "6.0.json"))
    assert set(d) == set(pkg.Loss.keys) | {'T', 'EbN0dB', 'SNRdB', 'rate', 'C', 'ShannonLimitdB'}   # loss.py:27,313-317
    assert d['fer'] == 0.25 and L.loss == {'T': 0}


def test_detectors_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg = pkg.Config(8, 1, 4, 1, 1, batch=2, generator_mode='sparc', alphabet='QPSK', channel_profile='uniform', device='cpu')
    y = torch.zeros(2, 4, 1, dtype=torch.complex64)
    with pytest.raises(_cabi.AmpsmError):
        pkg.BAMP(cfg)(torch.zeros(4, 8, dtype=torch.complex64), y, 1.0, torch.zeros(2, 8, 1, dtype=torch.complex64),
                      np.zeros(2, int), np.zeros(2, int))
    with pytest.raises(_cabi.AmpsmError):
        pkg.Loss(cfg)(y, y, y, np.zeros(2, int), np.zeros(2, int), 1)


def test_cabi_library_exports_every_declared_symbol():
    """The shared object loads (no GPU needed) and exports exactly what include/ampsm_b200.h declares."""
    header = open(os.path.join(ROOT, "include", "ampsm_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ampsm_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.EXPORTS)
    lib = _cabi.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in lib.ampsm_version()


def test_cabi_struct_layout_matches_header():
    import ctypes
    assert ctypes.sizeof(_cabi.Alphabet) == 4 + 4 * 16 + 4 + 8 * 16 * 2 or ctypes.sizeof(_cabi.Alphabet) == 328
    assert ctypes.sizeof(_cabi.Problem) == 16 * 4 + 8


def test_cabi_struct_layout_matches_a_c_compiler(tmp_path):
    """sizeof / offsetof of every struct of include/ampsm_b200.h as gcc lays them out against the ctypes mirrors."""
    import ctypes
    import subprocess
    fields = {"ampsm_alphabet": (_cabi.Alphabet, ["K", "gray", "re", "im"]),
              "ampsm_problem": (_cabi.Problem, ["n", "R", "max_iters", "decision", "kernel", "frame_base"]),
              "ampsm_gen": (_cabi.Gen, ["seed", "counter_base", "h_var", "Rr_root", "Rt_root", "real_roots", "rho_r", "rho_t"])}
    src = ['#include <stdio.h>', '#include <stddef.h>', '#include "ampsm_b200.h"', 'int main(void) {']
    for name, (_, fl) in fields.items():
        src.append(f'printf("{name} %zu", sizeof({name}));')
        for f in fl:
            src.append(f'printf(" %zu", offsetof({name}, {f}));')
        src.append('printf("\\n");')
    src.append('return 0; }')
    c = tmp_path / "layout.c"
    c.write_text("\n".join(src))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    for line in out.strip().splitlines():
        name, size, *offs = line.split()
        ct, fl = fields[name]
        assert ctypes.sizeof(ct) == int(size), name
        assert [getattr(ct, f).offset for f in fl] == [int(o) for o in offs], name


def test_device_frames_generator_is_consistent_on_cpu():
    """simulate.device_frames (the on-device input generator of the Monte-Carlo driver) with a CPU generator: labels and
    flat indices describe x as Data.generate_message does (data.py:88-90), y - H x is noise of variance Na/Nr/SNR, and
    the Kronecker option imposes the exponential correlation."""
    import torch
    from amp_sparc_spatialmodulation_b200.simulate import device_frames, exp_corr_root
    cfg = pkg.Config(64, 4, 32, 1, 1, batch=4096, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='uniform', device='cpu')
    gen = torch.Generator(device='cpu').manual_seed(5)
    snr = 10.0
    H, y, x, lab, idx = device_frames(cfg, 4096, snr, gen)
    F, N, L, M = 4096, cfg.N, cfg.L, cfg.M
    nz = x.reshape(-1).nonzero().reshape(-1)
    assert torch.equal(nz, idx) and idx.numel() == F * L                      # sorted flat positions, one per section
    assert bool(((idx % N) // M == torch.arange(L).repeat(F)).all())
    sym = torch.as_tensor(np.asarray(cfg.symbols)).to(torch.complex64)
    gray = np.asarray(cfg.gray)
    k = (x.reshape(-1)[idx][:, None] - sym[None, :]).abs().argmin(dim=1)
    assert np.array_equal(gray[k.numpy()], lab.numpy())
    noise = y - (H @ x.unsqueeze(-1)).squeeze(-1)
    sigma2 = (cfg.Na / cfg.Nr) / snr
    assert abs(float(noise.abs().pow(2).mean()) / sigma2 - 1) < 0.02
    assert abs(float(H.abs().pow(2).mean()) * cfg.Nr - 1) < 0.02
    Hk, *_ = device_frames(cfg, 4096, snr, gen, channel='kronecker', rho_t=0.7, rho_r=0.5)
    Rt = (Hk.mH @ Hk).mean(dim=0) / (Hk.abs().pow(2).sum(dim=1).mean())         # column correlation, unit diagonal
    assert abs(float(Rt[0, 1].real) - 0.7) < 0.05 and abs(float(Rt[0, 2].real) - 0.49) < 0.05
    root = exp_corr_root(8, 0.9, 'cpu').to(torch.complex128)
    i = torch.arange(8, dtype=torch.float64)
    assert float(((root @ root).real - 0.9 ** (i[:, None] - i[None, :]).abs()).abs().max()) < 1e-6


@pytest.mark.parametrize("trunc", ["trunc", "tail", "cyclic"])
def test_device_isi_frames_match_the_dense_block_toeplitz_matrix(trunc):
    """ISI frames of the sweep driver: y - H x with H rebuilt from the taps (matrix_from_taps, the layout of
    channel.py:56-72) is pure noise of variance Na/Nr/SNR, and the taps carry the power of channel.py:55."""
    import torch
    from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps
    from amp_sparc_spatialmodulation_b200.simulate import device_frames
    cfg = pkg.Config(12, 2, 6, 5, 3, batch=512, generator_mode='sparc', iterations=20, alphabet='QPSK',
                     channel_profile='exponential', channel_truncation=trunc, device='cpu')
    gen = torch.Generator(device='cpu').manual_seed(9)
    snr = 1e6
    taps, y, x, lab, idx = device_frames(cfg, 512, snr, gen)
    assert taps.shape == (512, 3, 6, 12) and y.shape == (512, cfg.n)
    H = matrix_from_taps(taps, cfg.Lin, cfg.Lout, trunc == 'cyclic')
    resid = y - (H @ x.unsqueeze(-1)).squeeze(-1)
    sigma2 = (cfg.Na / cfg.Nr) / snr
    assert float(resid.abs().max()) < 20 * np.sqrt(sigma2) + 1e-5
    assert torch.equal(x.reshape(-1).nonzero().reshape(-1), idx)
    pdp = np.exp(-np.arange(3)); pdp /= pdp.sum()
    power = taps.abs().pow(2).mean(dim=(0, 2, 3)).numpy()
    assert np.allclose(power, pdp * cfg.Lout / cfg.Nr / cfg.Lin, rtol=0.05)


def test_dense_materialises_lazy_conjugate_views():
    """_tensors.dense: what reaches a kernel pointer is plain memory holding the tensor's VALUES (conj / neg bits resolved)."""
    import torch
    from amp_sparc_spatialmodulation_b200._tensors import dense
    a = torch.randn(3, 5, dtype=torch.complex64)
    v = a.mH                                                   # conj-view, non-contiguous
    d = dense(v, 'cpu', torch.complex64)
    assert v.is_conj() and not d.is_conj() and d.is_contiguous() and torch.equal(d, a.conj().T.contiguous())
    w = a.conj()                                               # conj-view with contiguous strides: .contiguous() keeps the bit
    assert w.contiguous().is_conj() and not dense(w, 'cpu', torch.complex64).is_conj()
    assert np.array_equal(np.frombuffer(dense(w, 'cpu', torch.complex64).numpy().tobytes(), np.complex64),
                          a.conj().resolve_conj().numpy().ravel())
    assert dense(a, 'cpu', torch.complex64, -1, 1).shape == (15, 1)


def test_scamp_design_matrix_is_recognised_as_block_toeplitz():
    """Channel.generate_as_sparc (channel.py:76-96) builds A = sum_l kron(eye(Lout, Lin, -l) sqrt(W), h_l): block (r, c) = taps[r - c].
    The host-side structure test behind SCAMP's structured route must recover the taps bit-exactly for 'tail' and 'trunc'
    layouts, rebuild A from them, and reject a matrix that is not block-Toeplitz."""
    import torch
    from amp_sparc_spatialmodulation_b200.bamp import matrix_from_taps, taps_from_matrix
    for shape, trunc in (((64, 2, 8, 8, 3), 'tail'), ((48, 2, 6, 5, 2), 'tail'), ((64, 2, 8, 8, 3), 'trunc'), ((32, 4, 16, 6, 1), 'trunc')):
        cfg = pkg.Config(*shape, batch=2, generator_mode='sparc', iterations=20, alphabet='QPSK', channel_profile='exponential',
                         channel_truncation=trunc, device='cpu')
        np.random.seed(3)
        W, A = pkg.Channel(cfg).generate_as_sparc()
        st = taps_from_matrix(A, cfg)
        assert st is not None and st[1] is False
        taps = st[0]
        assert tuple(taps.shape) == (min(cfg.Lh, cfg.Lout), cfg.Nr, cfg.Nt)
        assert torch.equal(matrix_from_taps(taps, cfg.Lin, cfg.Lout), A)
        if cfg.Lin > 1:
            B = A.clone()
            B[-1, 0] = 1.0 + 0j                                    # an entry outside the band
            assert taps_from_matrix(B, cfg) is None
            C_ = A.clone()
            C_[cfg.Nr, cfg.Nt] += 1e-3                              # block (1, 1) no longer equals block (0, 0)
            assert taps_from_matrix(C_, cfg) is None
