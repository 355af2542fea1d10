"""Hard-decision scoring with the reference ``Loss`` surface (/root/reference/loss.py:8-349).

Same constructor, ``keys``, ``loss`` dict, ``__call__`` / ``error_rate`` / ``accumulate`` / ``average`` / ``export`` /
``dump``.  The decision (MAP, loss.py:282-302, or segmented, 223-250) and all error counts are computed on the GPU
by ``ampsm_loss_count`` (or fused into the detector kernels); this class only turns the integer counters into the
14 rates with the reference's own formulas (loss.py:116-178).  There is no CPU path.
"""
import json
import math

import numpy as np
import torch

from . import _cabi
from ._tensors import dense
from .config import Config


class Loss:
    keys = ['fer', 'nMSE', 'nMSEf', 'nMSEm', 'nMSEL', 'ver', 'verf', 'verm', 'verL', 'ber', 'iber', 'sber', 'ier', 'ser']

    def __init__(self, config: Config) -> None:
        self.config = config
        self.B, self.Nt, self.Na, self.Nr, self.Lin = config.B, config.Nt, config.Na, config.Nr, config.Lin
        self.Ns, self.sparsity = config.Ns, config.sparsity
        self.gray = config.gray
        self.symbols = config.symbols
        count = self.Lin * self.B * self.Na
        self._ibits = int(math.ceil(math.log2(count))) if count > 0 else 0     # loss.py:20
        self.ibits = config.index_bits
        self.sbits = config.symbol_bits
        self.rate = config.code_rate
        self.shannon_limt_dB = config.shannon_limit_dB                            # (sic) loss.py:24
        self.keys = list(Loss.keys)
        self.loss = {'T': 0}
        self.counters = None           # counters of the last call (dict), an addition to the reference surface

    # -- counters -> rates ---------------------------------------------------------------------------------
    def rates(self, c):
        """The 14 rates for a call that held c['frames'] frames (B = frames in loss.py:116-178)."""
        B, Na, Lin = c['frames'], self.Na, self.Lin
        Ns = B * Lin * Na
        iber_ = c['index_bit_err'] / Lin / B
        with np.errstate(divide='ignore', invalid='ignore'):
            iber = float(np.float64(iber_) / np.float64(self.ibits))
        if self.sbits != 0:
            sber_ = c['symbol_bit_err'] / Lin / B
            sber = sber_ / self.sbits / Na
        else:
            sber_, sber = 0., 0.
        return dict(fer=c['frame_err'] / B,
                    nMSE=c['sqerr'] / Ns, nMSEf=c['sqerr_first'] / Na / B, nMSEm=c['sqerr_mid'] / Na / B,
                    nMSEL=c['sqerr_last'] / Na / B,
                    ver=c['slot_err'] / Lin / B, verf=c['slot_err_first'] / B, verm=c['slot_err_mid'] / B,
                    verL=c['slot_err_last'] / B,
                    ber=(iber_ + sber_) / (Na * self.sbits + self.ibits), iber=iber, sber=sber,
                    ier=c['index_err'] / Ns, ser=c['symbol_err'] / Ns)

    def record(self, counters, iterations):
        """Store one call's result: same effect on ``self.loss`` as loss.py:58-65."""
        self.counters = counters
        self.loss['T'] = iterations
        r = self.rates(counters)
        for key in self.keys:
            if key in self.loss:
                self.loss[key] = np.append(self.loss[key], r[key])
            else:
                self.loss[key] = np.array(r[key])

    # -- reference surface ---------------------------------------------------------------------------------
    def __call__(self, xmap, xmmse, x, symbols, indices, iterations) -> None:
        self.record(self._count(xmap, xmmse, x, symbols, indices), iterations)

    def error_rate(self, xmap, xmmse, x, symbols=None, indices=None):
        r = self.rates(self._count(xmap, xmmse, x, symbols, indices))
        return tuple(r[k] for k in self.keys)

    def _count(self, xmap, xmmse, x, symbols, indices):
        if not torch.cuda.is_available():
            raise _cabi.AmpsmError("Loss needs a CUDA device: decisions and counters run in ampsm_loss_count (no CPU path)")
        lib = _cabi.lib()
        dev = xmap.device if xmap.is_cuda else torch.device('cuda', torch.cuda.current_device())
        N = self.Nt * self.Lin
        xm = dense(xmap, dev, torch.complex64, -1, N)
        xe = dense(xmmse, dev, torch.complex64, -1, N)
        xt = dense(x, dev, torch.complex64, -1, N)
        F = xt.shape[0]
        sym = torch.as_tensor(np.ascontiguousarray(symbols, dtype=np.int64)).to(dev)
        idx = torch.as_tensor(np.ascontiguousarray(indices, dtype=np.int64)).to(dev)
        counters = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            prob = _cabi.make_problem(self.config, F)
            alpha = _cabi.make_alphabet(self.config)
            rc = lib.ampsm_loss_count(prob, alpha, F, xm.data_ptr(), xe.data_ptr(), xt.data_ptr(), sym.data_ptr(),
                                      idx.data_ptr(), None, counters.data_ptr(), torch.cuda.current_stream(dev).cuda_stream)
            _cabi.check(rc, "ampsm_loss_count")
        return _cabi.counters_to_dict(counters.cpu().numpy())

    def export(self, SNRdB: float, EbN0dB: float, save_location: str) -> None:
        """Write ``<save_location>/<EbN0dB>.json`` with the reference's schema (loss.py:304-323)."""
        self.loss['EbN0dB'] = float(EbN0dB)
        self.loss['SNRdB'] = float(SNRdB)
        self.loss['rate'] = float(self.rate)
        self.loss['C'] = float(np.log2(1 + 10 ** (SNRdB / 10)))
        self.loss['ShannonLimitdB'] = float(self.shannon_limt_dB)

        def plain(v):
            a = np.asarray(v)
            return float(a) if a.ndim == 0 else [float(e) for e in a.ravel()]
        with open(f'{save_location}/{EbN0dB}.json', 'w', encoding='utf-8') as f:
            json.dump({k: plain(v) for k, v in self.loss.items()}, f, ensure_ascii=False, indent=6)
        self.loss = {'T': 0}

    def accumulate(self, other) -> None:
        self.loss['T'] += other.loss['T']
        for key in self.keys:
            if key in self.loss:
                self.loss[key] = self.loss[key] + other.loss[key]
            else:
                self.loss[key] = other.loss[key]

    def average(self, epochs: int) -> None:
        self.loss['T'] = self.loss['T'] / epochs
        for key in self.keys:
            self.loss[key] = np.array(self.loss[key]) / epochs

    def dump(self) -> None:
        self.loss = {}
