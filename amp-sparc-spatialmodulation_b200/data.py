"""Spatial-modulation message generator (an input generator, not part of the hot path).

Same surface and the same ``np.random`` call order as the reference ``Data`` (/root/reference/data.py:7-91):
``generate_message() -> (x, gray_labels, flat_indices)`` with ``x`` of shape (B, Nt*Lin, 1).
"""
import numpy as np
import torch

from .config import Config


class Data:
    def __init__(self, config: Config) -> None:
        self.B, self.Lin = config.B, config.Lin
        self.Nt, self.Na = config.Nt, config.Na
        self.Ns = config.Na
        self.device = config.device
        self.symbols = config.symbols
        self.gray = config.gray
        self.cardinality = len(self.symbols)
        self.dtype = torch.complex64 if config.is_complex else torch.float32
        self.npdtype = np.complex64 if config.is_complex else np.float32
        if config.mode == 'random':
            self._generator = self.random
        else:
            assert self.Nt % self.Na == 0, 'Na must divide Nt'
            self.L = self.Na * self.Lin
            self.M = self.Nt // self.Na
            self._generator = self.segmented

    def generate_message(self):
        x, labels, index = self._generator()
        return torch.tensor(x, device=self.device, dtype=self.dtype, requires_grad=False), labels, index

    def _finish(self, x, xgray):
        x = x.reshape(self.B, -1, 1)
        index = x.ravel().nonzero()[0]
        return x, xgray.ravel()[index], index

    def random(self):
        """Na distinct antennas out of Nt per time slot, one shared symbol (data.py:55-72)."""
        x = np.zeros((self.B, self.Lin, self.Nt), dtype=self.npdtype)
        xgray = np.zeros((self.B, self.Lin, self.Nt), dtype=int)
        for b in range(self.B):
            for t in range(self.Lin):
                where = np.random.choice(self.Nt, size=self.Na, replace=False)
                k = np.random.choice(self.cardinality)
                x[b, t, where] = self.symbols[k]
                xgray[b, t, where] = self.gray[k]
        return self._finish(x, xgray)

    def segmented(self):
        """One active antenna per section of M = Nt/Na with its own symbol (data.py:74-91)."""
        x = np.zeros((self.B, self.L, self.M), dtype=self.npdtype)
        xgray = np.zeros((self.B, self.L, self.M), dtype=int)
        for b in range(self.B):
            for sec in range(self.L):
                where = np.random.choice(self.M)
                k = np.random.choice(self.cardinality)
                x[b, sec, where] = self.symbols[k]
                xgray[b, sec, where] = self.gray[k]
        return self._finish(x, xgray)
