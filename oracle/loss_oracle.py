"""CPU oracle for the hard-decision / error-metric epilogue -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Vectorised numpy restatement of the reference ``Loss`` (/root/reference/loss.py):
  * ``map_decision``        -> loss.py:282-302 (mode 'sparc'; first maximum of Re(x conj(sym)) in complex128,
                               row-major over (antenna, symbol); NaN wins as in np.argmax)
  * ``segmented_decision``  -> loss.py:223-250 (strongest antenna by |x|, then nearest symbol, first minimum)
  * ``random_decision``     -> loss.py:252-280 (mode 'random': the Na strongest entries of every time slot, each
                               mapped to its nearest symbol)
  * ``error_counters``      -> the integer counts / squared-error sums behind loss.py:105-179
  * ``rates_from_counters`` -> the 14 rates of loss.py:27 formed exactly as loss.py:116-178 forms them

Pinned by tests/test_oracle_golden.py against the reference's own ``Loss`` outputs stored in tests/golden.
See oracle/amp_oracle.py for who may import this package.
"""
import math

import numpy as np

KEYS = ['fer', 'nMSE', 'nMSEf', 'nMSEm', 'nMSEL', 'ver', 'verf', 'verm', 'verL', 'ber', 'iber', 'sber', 'ier', 'ser']
COUNTER_KEYS = ['frames', 'frame_err', 'slot_err', 'slot_err_first', 'slot_err_mid', 'slot_err_last',
                'index_err', 'symbol_err', 'index_bit_err', 'symbol_bit_err', 'iters', 'nan_frames']
SQERR_KEYS = ['sqerr', 'sqerr_first', 'sqerr_mid', 'sqerr_last']


def map_decision(xmap, symbols, gray, M):
    """xmap: (..., N) complex64 -> (xhat (S, M) complex64, best antenna (S,), best symbol (S,)), S = sections."""
    xs = np.ascontiguousarray(xmap, dtype=np.complex64).reshape(-1, M)
    sym = np.asarray(symbols, dtype=np.complex128)
    with np.errstate(all='ignore'):
        metric = (xs.astype(np.complex128)[:, :, None] * sym.conj()[None, None, :]).real
    flat = metric.reshape(xs.shape[0], -1).argmax(axis=1)            # np.argmax: first max, NaN counts as max
    ant, k = np.divmod(flat, sym.shape[0])
    xhat = np.zeros_like(xs)
    xhat[np.arange(xs.shape[0]), ant] = sym[k]                        # rounds to complex64 (loss.py:297)
    return xhat, ant, k


def segmented_decision(xmap, symbols, gray, M):
    xs = np.ascontiguousarray(xmap, dtype=np.complex64).reshape(-1, M)
    sym = np.asarray(symbols, dtype=np.complex128)
    ant = np.abs(xs).argsort(axis=1)[:, -1]                           # loss.py:236
    picked = xs[np.arange(xs.shape[0]), ant]
    with np.errstate(all='ignore'):
        dist = np.abs(picked[:, None] - sym[None, :])                 # complex128 distance (loss.py:240)
    k = np.zeros(xs.shape[0], dtype=np.int64)
    best = np.full(xs.shape[0], np.inf)
    for i in range(sym.shape[0]):                                     # strict '<' keeps the first minimum; NaN never wins
        better = dist[:, i] < best
        k = np.where(better, i, k)
        best = np.where(better, dist[:, i], best)
    won = np.isfinite(best) | (best < np.inf)
    xhat = np.zeros_like(xs)
    rows = np.arange(xs.shape[0])[won]
    xhat[rows, ant[won]] = sym[k[won]]
    return xhat, ant, k


def random_decision(xmap, symbols, gray, Nt, Na):
    """loss.py:252-280 per time slot of Nt entries: positions of the Na largest |x| (``argsort()[-Na:]``), each decided
    to the nearest symbol (first minimum of the complex128 distance).  Returns (xhat (S, Nt) complex64, xgray (S, Nt))."""
    xs = np.ascontiguousarray(xmap, dtype=np.complex64).reshape(-1, Nt)
    sym = np.asarray(symbols, dtype=np.complex128)
    gray = np.asarray(gray, dtype=np.int64)
    xhat = np.zeros_like(xs)
    xgray = np.zeros(xs.shape, dtype=np.int64)
    top = np.abs(xs).argsort(axis=1)[:, -Na:]                          # loss.py:266
    rows = np.arange(xs.shape[0])[:, None]
    picked = xs[rows, top]                                            # (S, Na) complex64
    with np.errstate(all='ignore'):
        dist = np.abs(picked[:, :, None].astype(np.complex128) - sym[None, None, :])
    k = np.zeros(top.shape, dtype=np.int64)
    best = np.full(top.shape, np.inf)
    for i in range(sym.shape[0]):                                     # strict '<' keeps the first minimum (loss.py:271)
        better = dist[:, :, i] < best
        k = np.where(better, i, k)
        best = np.where(better, dist[:, :, i], best)
    won = best < np.inf
    r2 = np.broadcast_to(rows, top.shape)
    xhat[r2[won], top[won]] = sym[k[won]]
    xgray[r2[won], top[won]] = gray[k[won]]
    return xhat, xgray


def error_counters(xmap, xmmse, x, sym_true, idx_true, symbols, gray, dims, iters=None, decision='sparc',
                   index_bits_kept=None):
    """Counts behind the 14 metrics for a call holding F frames (loss.py:67-179).

    dims: dict with Nt, Na, Lin; ``index_bits_kept`` defaults to ceil(log2(Lin*F*Na)) (loss.py:20 with B=F).
    Returns a dict of python ints / floats keyed by COUNTER_KEYS + SQERR_KEYS.
    """
    Nt, Na, Lin = dims['Nt'], dims['Na'], dims['Lin']
    M = Nt // Na
    xmap = np.asarray(xmap).reshape(-1, Lin, Nt)
    xmmse = np.asarray(xmmse, dtype=np.complex64).reshape(-1, Lin, Nt)
    x = np.asarray(x, dtype=np.complex64).reshape(-1, Lin, Nt)
    F = x.shape[0]
    gray = np.asarray(gray, dtype=np.int64)
    if decision == 'random':
        xhat_sec, xgray_r = random_decision(xmap, symbols, gray, Nt, Na)
    else:
        decide = map_decision if decision == 'sparc' else segmented_decision
        xhat_sec, ant, k = decide(xmap, symbols, gray, M)
    xhat = xhat_sec.reshape(-1, Lin, Nt)

    d = (xmmse - x)
    se = d.real.astype(np.float64) ** 2 + d.imag.astype(np.float64) ** 2          # (F, Lin, Nt)
    mism = (xhat != x)                                                # value compare in complex64 (loss.py:133,150)
    slot_bad = mism.any(axis=-1)                                      # (F, Lin)
    mid = Lin // 2

    flat = xhat_sec.ravel()
    idx_hat = np.sort(flat.nonzero()[0])                              # loss.py:300
    if decision == 'random':
        xgray = xgray_r
    else:
        xgray = np.zeros(xhat_sec.shape, dtype=np.int64)
        xgray[np.arange(xhat_sec.shape[0]), ant] = gray[k]
    sym_hat = xgray.ravel()[idx_hat]
    idx_true = np.asarray(idx_true, dtype=np.int64)
    sym_true = np.asarray(sym_true, dtype=np.int64)
    if index_bits_kept is None:
        index_bits_kept = int(math.ceil(math.log2(Lin * F * Na)))
    sbits = int(np.log2(len(symbols)))

    def popcount_low(v, bits):
        if bits <= 0:
            return 0
        v = v.astype(np.int64) & ((1 << bits) - 1)
        return int(sum(int(((v >> b) & 1).sum()) for b in range(bits)))

    out = dict(
        frames=F,
        frame_err=int(mism.reshape(F, -1).any(axis=1).sum()),
        slot_err=int(slot_bad.sum()),
        slot_err_first=int(slot_bad[:, 0].sum()),
        slot_err_mid=int(slot_bad[:, mid].sum()),
        slot_err_last=int(slot_bad[:, -1].sum()),
        index_err=int(np.count_nonzero(idx_hat - idx_true)),
        symbol_err=int(np.count_nonzero(sym_hat - sym_true)),
        index_bit_err=popcount_low(np.bitwise_xor(idx_hat, idx_true), index_bits_kept),
        symbol_bit_err=popcount_low(np.bitwise_xor(sym_hat, sym_true), sbits),
        iters=int(np.sum(iters)) if iters is not None else 0,
        nan_frames=int(np.isnan(np.asarray(xmap).reshape(F, -1).real).any(axis=1).sum()
                       + 0),
        sqerr=float(se.sum()),
        sqerr_first=float(se[:, 0].sum()),
        sqerr_mid=float(se[:, mid].sum()),
        sqerr_last=float(se[:, -1].sum()),
    )
    return out


def rates_from_counters(c, dims, index_bits, symbol_bits):
    """The 14 rates of loss.py:27 from the counters, with B = c['frames'] (formulas loss.py:116-178)."""
    Na, Lin = dims['Na'], dims['Lin']
    B = c['frames']
    Ns = B * Lin * Na
    iber_ = c['index_bit_err'] / Lin / B
    iber = iber_ / index_bits if index_bits != 0 else float('nan')
    if symbol_bits != 0:
        sber_ = c['symbol_bit_err'] / Lin / B
        sber = sber_ / symbol_bits / Na
    else:
        sber_, sber = 0., 0.
    return dict(
        fer=c['frame_err'] / B,
        nMSE=c['sqerr'] / Ns, nMSEf=c['sqerr_first'] / Na / B, nMSEm=c['sqerr_mid'] / Na / B,
        nMSEL=c['sqerr_last'] / Na / B,
        ver=c['slot_err'] / Lin / B, verf=c['slot_err_first'] / B, verm=c['slot_err_mid'] / B,
        verL=c['slot_err_last'] / B,
        ber=(iber_ + sber_) / (Na * symbol_bits + index_bits),
        iber=iber, sber=sber,
        ier=c['index_err'] / Ns, ser=c['symbol_err'] / Ns,
    )
