"""Generate the golden fixtures in this directory by running the REFERENCE itself (build container only).

    python tests/golden/make_golden.py [name ...]   # needs /root/reference (read-only) and CPU torch/numpy

The reference ships no tests, seeds or golden vectors, so the pins are manufactured here:
fixed-seed inputs drawn by the reference's own ``Channel`` / ``Data`` (numpy + torch global RNGs), detectors run
once per frame with ``batch=1`` (the reference's real operating mode), per-iteration trajectories captured by
stepping ``amp.layers[t](T)`` by hand, final ``xmap`` / ``xmmse`` / exit iteration and the ``Loss.loss`` dicts.
Nothing here is imported at test time; the tests only read the ``.npz`` files this script writes.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("AMPSM_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
import bamp as ref_bamp      # noqa: E402
import scamp as ref_scamp    # noqa: E402
import vamp as ref_vamp      # noqa: E402
import vamp2 as ref_vamp2    # noqa: E402
from channel import Channel  # noqa: E402
from config import Config    # noqa: E402
from data import Data        # noqa: E402
from loss import Loss        # noqa: E402
from shrink import Shrink    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
KEYS = ['fer', 'nMSE', 'nMSEf', 'nMSEm', 'nMSEL', 'ver', 'verf', 'verm', 'verL', 'ber', 'iber', 'sber', 'ier', 'ser']
torch.set_grad_enabled(False)


def cfg(Nt, Na, Nr, Lin, Lh, alphabet, trunc='trunc', mode='sparc', iters=20, batch=1):
    return Config(Nt, Na, Nr, Lin, Lh, batch=batch, generator_mode=mode, iterations=iters, alphabet=alphabet,
                  channel_profile='uniform', channel_truncation=trunc, device='cpu')


def loss_vec(L):
    return np.array([float(np.asarray(L.loss[k]).reshape(-1)[0]) for k in KEYS], dtype=np.float64)


def mse_of(xmmse, x):
    return float((xmmse - x).abs().pow(2).mean())


def batch_loss(config_args, frames, xmap, xmmse, x, sym, idx, N):
    """The reference's own Loss evaluated on the stacked frames with B = number of frames."""
    c = cfg(*config_args[0], **dict(config_args[1], batch=frames))
    L = Loss(c)
    gidx = np.concatenate([i + f * N for f, i in enumerate(idx)])
    L(torch.tensor(np.stack(xmap)).reshape(frames, N, 1), torch.tensor(np.stack(xmmse)).reshape(frames, N, 1),
      torch.tensor(np.stack(x)).reshape(frames, N, 1), np.concatenate(sym), gidx, 0)
    return loss_vec(L), gidx


def run_bamp(name, args, kwargs, snrs_db, frames, seed, matrix='channel', taps_only=False):
    c = cfg(*args, **kwargs)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ch, da, amp = Channel(c), Data(c), ref_bamp.BAMP(c)
    N, n, T = c.Nt * c.Lin, c.Nr * c.Lout, c.N_Layers
    out = {k: [] for k in ('H', 'y', 'x', 'sym', 'idx', 'sigma2', 'snr_db', 'xmap', 'xmmse', 'var', 'iters', 'tau', 'varm',
                           'mse', 'loss')}
    for snr_db in snrs_db:
        snr = 10 ** (snr_db / 10)
        for _ in range(frames):
            H = ch.generate_channel() if matrix == 'channel' else ch.generate_as_sparc()[1]
            x, s, i = da.generate_message()
            y = H @ x + ch.awgn(snr)
            tr = ref_bamp.Tracker(x, y, H, amp.E / snr)
            tau, varm, mse = np.full(T, np.nan), np.full(T, np.nan), np.full(T, np.nan)
            for t, layer in enumerate(amp.layers):
                prev = tr.var
                layer(tr)
                tau[t:] = float((1 / (tr.abs2T @ (1 / tr.u))).real.mean())
                varm[t:] = float(tr.var.mean())
                mse[t:] = mse_of(tr.xmmse, x)
                if torch.allclose(tr.var, prev):
                    break
            L = amp(H, y, snr, x, s, i)                      # the stock forward; must agree with the stepping
            assert L.loss['T'] == t + 1
            Hstore = H.numpy()
            if taps_only:
                # large ISI matrices: keep only the first block column (the Lh tap matrices); the reference's H is
                # exactly the block-Toeplitz matrix of these taps, checked here before anything is written
                Hn, (Nr, Nt, Lin, Lout, Lh) = H.numpy(), (c.Nr, c.Nt, c.Lin, c.Lout, c.Lh)
                Hstore = np.stack([Hn[l * Nr:(l + 1) * Nr, :Nt] for l in range(Lh)])
                re = np.zeros_like(Hn)
                for io in range(Lout):
                    for ji in range(Lin):
                        d = (io - ji) % Lin if c.trunc == 'cyclic' and matrix == 'channel' else io - ji
                        if 0 <= d < Lh:
                            re[io * Nr:(io + 1) * Nr, ji * Nt:(ji + 1) * Nt] = Hstore[d]
                assert np.array_equal(re, Hn)
            for k, v in (('H', Hstore), ('y', y.numpy().reshape(n)), ('x', x.numpy().reshape(N)), ('sym', s),
                         ('idx', i), ('sigma2', amp.E / snr), ('snr_db', snr_db), ('xmap', tr.xmap.numpy().reshape(N)),
                         ('xmmse', tr.xmmse.numpy().reshape(N)), ('var', tr.var.numpy().reshape(N)), ('iters', t + 1),
                         ('tau', tau), ('varm', varm), ('mse', mse), ('loss', loss_vec(L))):
                out[k].append(v)
    F = len(out['H'])
    bl, gidx = (None, None)
    if c.mode == 'sparc':
        bl, gidx = batch_loss((args, kwargs), F, out['xmap'], out['xmmse'], out['x'], out['sym'], out['idx'], N)
    save(name, out, bl, gidx, dict(args=args, kwargs=kwargs, matrix=matrix, seed=seed, alg='bamp'))


def exp_corr_root(m, rho):
    """Hermitian square root of the exponential correlation matrix R[i, j] = rho^|i-j| (float64)."""
    i = np.arange(m)
    w, V = np.linalg.eigh(rho ** np.abs(i[:, None] - i[None, :]).astype(np.float64))
    return (V * np.sqrt(np.clip(w, 0, None))) @ V.T


def run_vamp(name, args, kwargs, snrs_db, frames, seed, double=False, kronecker=None):
    """kronecker = (rho_r, rho_t): BASELINE config 5 -- the channel is H = Rr^(1/2) G Rt^(1/2) with G the reference's own
    i.i.d. draw (Channel.generate_channel, CN(0, 1/Nr)) and exponential correlation on both sides; the reference has no
    correlated generator, its VAMP is fed torch.linalg.svd(H) exactly as vamp_model.py:56-61 does."""
    c = cfg(*args, **kwargs)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ch, da, amp = Channel(c), Data(c), ref_vamp.VAMP(c)
    N, n, T = c.Nt * c.Lin, c.Nr * c.Lout, c.N_Layers
    out = {k: [] for k in ('A', 'U', 's', 'Vh', 'y', 'x', 'sym', 'idx', 'sigma2', 'snr_db', 'xmap', 'xmmse', 'var', 'iters',
                           'tau', 'varm', 'mse', 'sigma2t', 'loss')}
    for snr_db in snrs_db:
        snr = 10 ** (snr_db / 10)
        for _ in range(frames):
            if kronecker:
                G = ch.generate_channel().numpy().astype(np.complex128)
                A = torch.tensor(exp_corr_root(n, kronecker[0]) @ G @ exp_corr_root(N, kronecker[1]), dtype=torch.complex64)
            else:
                _, A = ch.generate_as_sparc()
            U, s, Vh = torch.linalg.svd(A, full_matrices=False)
            x, sym, i = da.generate_message()
            y = A @ x + ch.awgn(snr)
            if double:
                Ui, si, Vhi, yi, xi = U.to(torch.complex128), s.to(torch.float64), Vh.to(torch.complex128), \
                    y.to(torch.complex128), x.to(torch.complex128)
            else:
                Ui, si, Vhi, yi, xi = U, s, Vh, y, x
            tr = ref_vamp.Tracker(Ui, si, Vhi, yi, xi, amp.E / snr, amp.sparsity)
            tau, varm, mse, s2t = (np.full(T, np.nan) for _ in range(4))
            for t, layer in enumerate(amp.layers):
                prev = tr.var
                layer(tr)
                s2t[t:] = float(tr.sigma2_tilde)
                varm[t:] = float(tr.var.mean())
                mse[t:] = mse_of(tr.xmmse, x)
                if torch.allclose(tr.var, prev):
                    break
            L = amp(Ui, si, Vhi, yi, snr, xi, sym, i)
            assert L.loss['T'] == t + 1
            for k, v in (('A', A.numpy()), ('U', U.numpy()), ('s', s.numpy()), ('Vh', Vh.numpy()), ('y', y.numpy().reshape(n)),
                         ('x', x.numpy().reshape(N)), ('sym', sym), ('idx', i), ('sigma2', amp.E / snr), ('snr_db', snr_db),
                         ('xmap', tr.r.numpy().reshape(N)), ('xmmse', tr.xmmse.numpy().reshape(N)),
                         ('var', tr.var.numpy().reshape(N)), ('iters', t + 1), ('tau', tau), ('varm', varm), ('mse', mse),
                         ('sigma2t', s2t), ('loss', loss_vec(L))):
                out[k].append(v)
    F = len(out['U'])
    xm32 = [np.asarray(v).astype(np.complex64) for v in out['xmap']]
    bl, gidx = batch_loss((args, kwargs), F, xm32, out['xmmse'], out['x'], out['sym'], out['idx'], N)
    save(name, out, bl, gidx, dict(args=args, kwargs=kwargs, seed=seed, alg='vamp', double=double, kronecker=kronecker))


def run_vamp2(name, args, kwargs, snrs_db, frames, seed, damping):
    """vamp2.py (the damped direct form of Rangan's VAMP that no driver of the reference imports): the reference's own class
    stepped layer by layer on the reference's own draws, fed torch.linalg.svd like its sibling (vamp_model.py:56-61)."""
    c = cfg(*args, **kwargs)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ch, da, amp = Channel(c), Data(c), ref_vamp2.VAMP(c, damping)
    N, n, T = c.Nt * c.Lin, c.Nr * c.Lout, c.N_Layers
    out = {k: [] for k in ('A', 'U', 's', 'Vh', 'y', 'x', 'sym', 'idx', 'sigma2', 'snr_db', 'xmap', 'xmmse', 'var', 'iters',
                           'gamma', 'varm', 'mse', 'loss')}
    for snr_db in snrs_db:
        snr = 10 ** (snr_db / 10)
        for _ in range(frames):
            _, A = ch.generate_as_sparc()
            U, s, Vh = torch.linalg.svd(A, full_matrices=False)
            x, sym, i = da.generate_message()
            y = A @ x + ch.awgn(snr)
            tr = ref_vamp2.Tracker(U, s, Vh, y, x, amp.E / snr)
            gam, varm, mse = (np.full(T, np.nan) for _ in range(3))
            for t, layer in enumerate(amp.layers):
                prev = tr.var
                layer(tr)
                gam[t:] = float(tr.gamma)
                varm[t:] = float(tr.var.mean())
                mse[t:] = mse_of(tr.xmmse, x)
                if torch.allclose(tr.var, prev):
                    break
            L = amp(U, s, Vh, y, snr, x, sym, i)
            assert L.loss['T'] == t + 1
            for k, v in (('A', A.numpy()), ('U', U.numpy()), ('s', s.numpy()), ('Vh', Vh.numpy()), ('y', y.numpy().reshape(n)),
                         ('x', x.numpy().reshape(N)), ('sym', sym), ('idx', i), ('sigma2', amp.E / snr), ('snr_db', snr_db),
                         ('xmap', tr.r.numpy().reshape(N)), ('xmmse', tr.xmmse.numpy().reshape(N)),
                         ('var', tr.var.numpy().reshape(N)), ('iters', t + 1), ('gamma', gam), ('varm', varm), ('mse', mse),
                         ('loss', loss_vec(L))):
                out[k].append(v)
    F = len(out['U'])
    bl, gidx = batch_loss((args, kwargs), F, out['xmap'], out['xmmse'], out['x'], out['sym'], out['idx'], N)
    save(name, out, bl, gidx, dict(args=args, kwargs=kwargs, seed=seed, alg='vamp2', damping=damping))


def run_scamp(name, args, kwargs, snrs_db, frames, seed, res):
    c = cfg(*args, **kwargs)
    np.random.seed(seed)
    torch.manual_seed(seed)
    ch, da, amp = Channel(c), Data(c), ref_scamp.SCAMP(c)
    N, n, T = c.Nt * c.Lin, c.Nr * c.Lout, c.N_Layers
    out = {k: [] for k in ('W', 'A', 'a_of_frame', 'y', 'x', 'sym', 'idx', 'sigma2', 'snr_db', 'xmap', 'xmmse', 'psi', 'iters',
                           'tau', 'psim', 'mse', 'loss')}
    for snr_db in snrs_db:
        snr = 10 ** (snr_db / 10)
        for f in range(frames):
            if f % res == 0:
                W, A = ch.generate_as_sparc()
                out['W'].append(W.numpy())
                out['A'].append(A.numpy())
            x, sym, i = da.generate_message()
            y = A @ x + ch.awgn(snr)
            tr = ref_scamp.Tracker(W, A, y, amp.E / snr, x)
            lay = amp.layers[0]
            tau, psim, mse = (np.full(T, np.nan) for _ in range(3))
            for t, layer in enumerate(amp.layers):
                prev = tr.psi
                layer(tr)
                tau[t:] = float((lay.L / (tr.W.T @ (1 / tr.phi)) / lay.Mr).mean())
                psim[t:] = float(tr.psi.mean())
                mse[t:] = mse_of(tr.xmmse, x)
                if torch.allclose(tr.psi, prev):
                    break
            L = amp(W, A, y, snr, x, sym, i)
            assert L.loss['T'] == t + 1
            for k, v in (('a_of_frame', len(out['A']) - 1), ('y', y.numpy().reshape(n)), ('x', x.numpy().reshape(N)),
                         ('sym', sym), ('idx', i), ('sigma2', amp.E / snr), ('snr_db', snr_db),
                         ('xmap', tr.xmap.numpy().reshape(N)), ('xmmse', tr.xmmse.numpy().reshape(N)),
                         ('psi', tr.psi.numpy().reshape(-1)), ('iters', t + 1), ('tau', tau), ('psim', psim), ('mse', mse),
                         ('loss', loss_vec(L))):
                out[k].append(v)
    F = len(out['y'])
    bl, gidx = batch_loss((args, kwargs), F, out['xmap'], out['xmmse'], out['x'], out['sym'], out['idx'], N)
    save(name, out, bl, gidx, dict(args=args, kwargs=kwargs, seed=seed, alg='scamp', res=res))


def run_loss_only(name, args, kwargs, frames, seed):
    """Reference Loss on synthetic estimates with B>1: pins decisions, tie-breaks and the index-bit truncation."""
    c = cfg(*args, **dict(kwargs, batch=frames))
    rng = np.random.RandomState(seed)
    np.random.seed(seed)
    da = Data(c)
    x, sym, idx = da.generate_message()
    N = c.Nt * c.Lin
    noise = (rng.normal(size=(frames, N, 1)) + 1j * rng.normal(size=(frames, N, 1))) * 0.35
    xmap = (x.numpy() + noise).astype(np.complex64)
    xmap[0, :, 0] = 0                                   # exact ties: the first (antenna, symbol) must win
    xmmse = (x.numpy() * 0.9 + 0.1 * noise).astype(np.complex64)
    L = Loss(c)
    L(torch.tensor(xmap), torch.tensor(xmmse), x, sym, idx, 0)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), xmap=xmap.reshape(frames, N), xmmse=xmmse.reshape(frames, N),
                        x=x.numpy().reshape(frames, N), sym=sym, idx=idx, loss=loss_vec(L),
                        meta=np.array(repr(dict(args=args, kwargs=kwargs, seed=seed, alg='loss'))))
    print(name, 'loss', dict(zip(KEYS, loss_vec(L).round(5))))


def run_shrink(name, seed):
    """Reference Shrink denoisers (shrink.py:58-157) on fixed-seed inputs: 'bayes' (complex QPSK prior) and
    'shrinkOOK' / 'sw_shrinkOOK' (OOK)."""
    rng = np.random.RandomState(seed)
    out = {}
    cq = cfg(16, 2, 8, 1, 1, 'QPSK', mode='random', batch=6)
    co = cfg(16, 2, 8, 1, 1, 'OOK', mode='segmented', batch=6)
    N = 16
    for tag, c in (('q', cq), ('o', co)):
        x, _, _ = Data(c).generate_message()
        r = (x.numpy() + (rng.normal(size=(6, N, 1)) + 1j * rng.normal(size=(6, N, 1))) * 0.3).astype(np.complex64)
        cov = (0.05 + 0.4 * rng.uniform(size=(6, N, 1))).astype(np.float32)
        out[f'r_{tag}'], out[f'cov_{tag}'] = r.reshape(6, N), cov.reshape(6, N)
    rq, covq = torch.tensor(out['r_q']).reshape(6, N, 1), torch.tensor(out['cov_q']).reshape(6, N, 1)
    ro, covo = torch.tensor(out['r_o']).reshape(6, N, 1), torch.tensor(out['cov_o']).reshape(6, N, 1)
    out['bayes'] = Shrink(cq, 'bayes')(rq.clone(), covq.clone()).numpy().reshape(6, N)
    # 'shrink' raises UnboundLocalError in the reference (d0 is used inside the tuple assignment that defines it,
    # shrink.py:113) and 'lasso' AttributeError (self.lmda, shrink.py:135): nothing to pin for those two.
    e, d = Shrink(co, 'shrinkOOK')(ro.clone(), covo.clone())
    out['ook_exp'], out['ook_dxdr'] = e.numpy().reshape(6, N), np.float32(d.numpy())
    e, v = Shrink(co, 'shrinkOOK').sw_shrinkOOK(ro.clone(), covo.clone())
    out['sw_exp'], out['sw_var'] = e.numpy().reshape(6, N), v.numpy().reshape(6, N)
    np.savez_compressed(os.path.join(HERE, name + '.npz'), **out)
    print(name, {k: v.shape for k, v in out.items()})


def save(name, out, batch_loss_vec, gidx, meta):
    arrays = {}
    for k, v in out.items():
        if k in ('sym', 'idx'):
            arrays[k] = np.stack(v).astype(np.int64)
        else:
            arrays[k] = np.stack([np.asarray(e) for e in v]) if len(v) else np.zeros(0)
    if batch_loss_vec is not None:
        arrays['batch_loss'] = batch_loss_vec
        arrays['batch_idx'] = gidx
    arrays['meta'] = np.array(repr(meta))
    path = os.path.join(HERE, name + '.npz')
    np.savez_compressed(path, **arrays)
    it = arrays['iters']
    print(f"{name}: frames={len(it)} mean T={it.mean():.2f} nan_frames={int(np.isnan(arrays['xmmse'].real).any(axis=1).sum())}"
          f" size={os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    only = set(sys.argv[1:])          # optional: fixture names to (re)generate; default all

    def want(name):
        return not only or name in only

    # C1: BAMP 8x4 QPSK, one active antenna (L=1)
    if want('bamp_c1'):
        run_bamp('bamp_c1', (8, 1, 4, 1, 1, 'QPSK'), {}, [0, 10, 20], 48, seed=0)
    # C2: BAMP 64x32 16-QAM (L=1) -- the headline config
    if want('bamp_c2'):
        run_bamp('bamp_c2', (64, 1, 32, 1, 1, '16QAM'), {}, [5, 10, 15, 20], 12, seed=0)
    # multi-section ISI frame, matrix drawn as in bamp_model.py:56 (generate_as_sparc)
    if want('bamp_isi'):
        run_bamp('bamp_isi', (16, 2, 8, 3, 2, 'QPSK'), dict(trunc='tail'), [2, 8], 12, seed=1, matrix='sparc')
    # structured-operator fixtures (taps only): cyclic and truncated convolution; a frame large enough for the four-slot tiles
    if want('bamp_isi_cyc'):
        run_bamp('bamp_isi_cyc', (32, 2, 12, 8, 3, 'QPSK'), dict(trunc='cyclic'), [4, 9], 6, seed=21, taps_only=True)
    if want('bamp_isi_trunc'):
        run_bamp('bamp_isi_trunc', (32, 2, 12, 8, 3, 'QPSK'), dict(trunc='trunc'), [4, 9], 6, seed=22, matrix='sparc', taps_only=True)
    if want('bamp_isi_big'):
        run_bamp('bamp_isi_big', (64, 4, 24, 24, 3, 'QPSK'), dict(trunc='tail'), [3, 7], 3, seed=23, taps_only=True)
    # 'segmented' decision rule (B=1 only in the reference)
    if want('bamp_seg'):
        run_bamp('bamp_seg', (16, 2, 8, 3, 2, '8PSK'), dict(trunc='tail', mode='segmented'), [4, 10], 8, seed=2, matrix='sparc')
    # C3: VAMP 128x64, Na=4, QPSK; complex64 and the complex128-input variant
    if want('vamp_c3'):
        run_vamp('vamp_c3', (128, 4, 64, 1, 1, 'QPSK'), {}, [0, 4], 3, seed=3)
    if want('vamp_c3_c128'):
        run_vamp('vamp_c3_c128', (128, 4, 64, 1, 1, 'QPSK'), {}, [0, 4], 2, seed=3, double=True)
    if want('vamp_isi'):
        run_vamp('vamp_isi', (16, 2, 8, 3, 2, 'QPSK'), dict(trunc='tail'), [4, 10], 8, seed=4)
    # VAMP at the headline 64x32 shapes (the register-resident kernel): 16-QAM one active antenna; QPSK, 4 sections of 16
    if want('vamp_c2'):
        run_vamp('vamp_c2', (64, 1, 32, 1, 1, '16QAM'), {}, [5, 10, 15, 20], 4, seed=8)
    if want('vamp_c2_na4'):
        run_vamp('vamp_c2_na4', (64, 4, 32, 1, 1, 'QPSK'), {}, [2, 8], 4, seed=9)
    # C5: VAMP on Kronecker-correlated (ill-conditioned) channels, rho = 0.7 and 0.9 on both sides, 64 x 32 QPSK
    if want('vamp_c5_rho07'):
        run_vamp('vamp_c5_rho07', (64, 1, 32, 1, 1, 'QPSK'), {}, [6, 12], 6, seed=31, kronecker=(0.7, 0.7))
    if want('vamp_c5_rho09'):
        run_vamp('vamp_c5_rho09', (64, 1, 32, 1, 1, 'QPSK'), {}, [10, 18], 6, seed=32, kronecker=(0.9, 0.9))
    # vamp2.py: the damped direct form (undamped and damping 0.97, the layer's own default)
    if want('vamp2_d100'):
        run_vamp2('vamp2_d100', (64, 4, 32, 1, 1, 'QPSK'), {}, [4, 10], 4, seed=41, damping=1.0)
    if want('vamp2_d097'):
        run_vamp2('vamp2_d097', (32, 2, 16, 1, 1, '16QAM'), {}, [8, 16], 4, seed=42, damping=0.97)
    # 'random' mode (i.i.d. prior, random_denoiser bamp.py:79-97, random_decision loss.py:252-280; B=1 only in the reference)
    if want('bamp_random'):
        run_bamp('bamp_random', (32, 4, 16, 1, 1, 'QPSK'), dict(mode='random'), [0, 6, 12], 8, seed=10)
    if want('bamp_random_isi'):
        run_bamp('bamp_random_isi', (16, 2, 8, 3, 2, '16QAM'), dict(trunc='tail', mode='random'), [4, 12], 6, seed=11)
    # SCAMP: small coupled instance, design matrix shared by groups of 4 frames (res=4)
    if want('scamp_small'):
        run_scamp('scamp_small', (32, 2, 8, 8, 3, 'QPSK'), dict(trunc='tail'), [4, 8], 8, seed=5, res=4)
    if want('shrink'):
        run_shrink('shrink', seed=12)
    # Loss alone at B>1
    if want('loss_qpsk'):
        run_loss_only('loss_qpsk', (16, 2, 8, 3, 2, 'QPSK'), dict(trunc='tail'), 24, seed=6)
    if want('loss_16qam'):
        run_loss_only('loss_16qam', (64, 1, 32, 1, 1, '16QAM'), {}, 64, seed=7)
