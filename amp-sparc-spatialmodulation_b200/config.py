"""Parameter bag for the spatial-modulation detector hot path.

Mirrors the constructor signature and every derived attribute of the reference ``Config``
(/root/reference/config.py:4-157) so that code written against the reference keeps working when only the
detector object is swapped.  Quirks that results-parity depends on are reproduced on purpose and are
flagged ``QUIRK`` below.
"""
import math

import numpy as np

# alphabet -> (constellation points before normalisation, Gray labels).
# QUIRK (config.py:112): the reference's 16QAM list holds -1+3j twice (entries 13 and 14, Gray 3 and 6) and has
# no 1-3j; the power normalisation is taken over this very list.  Kept verbatim: decisions, labels and the
# denoiser prior all depend on it.
_ALPHABETS = {
    "OOK": ([1], [1]),
    "BPSK": ([-1, 1], [0, 1]),
    "4ASK": ([-3, -1, 1, 3], [0, 1, 3, 2]),
    "QPSK": ([1 + 0j, 0 + 1j, -1 + 0j, 0 - 1j], [0, 1, 3, 2]),
    "8PSK": (None, [0, 1, 3, 2, 6, 7, 5, 4]),
    "16PSK": (None, [0, 1, 3, 2, 6, 7, 5, 4, 12, 13, 15, 14, 10, 11, 9, 8]),
    "16QAM": (
        [1 + 1j, 1 - 1j, -1 + 1j, -1 - 1j, 3 + 1j, 3 - 1j, -3 + 1j, -3 - 1j,
         3 + 3j, 3 - 3j, -3 + 3j, -3 - 3j, 1 + 3j, -1 + 3j, -1 + 3j, -1 - 3j],
        [0, 1, 13, 7, 8, 9, 2, 15, 12, 11, 5, 10, 14, 3, 6, 4],
    ),
}
_FORCES_COMPLEX = {"QPSK", "8PSK", "16PSK", "16QAM"}
_PROFILES = ("exponential", "uniform", "random")
_TRUNCATIONS = ("trunc", "tail", "cyclic")
_MODES = ("segmented", "random", "sparc")


def _constellation(alphabet):
    points, gray = _ALPHABETS[alphabet]
    if points is None:  # PSK rings, config.py:101/107
        order = len(gray)
        points = [np.exp((2 * np.pi * 1j / order) * k) for k in range(order)]
    return list(points), list(gray)


class Config:
    def __init__(self,
                 N_transmit_antenna: int,
                 N_active_antenna: int,
                 N_receive_antenna: int,
                 block_length: int,
                 channel_length: int,
                 batch: int = 100,
                 generator_mode: str = 'random',
                 iterations: int = 20,
                 alphabet: str = 'OOK',
                 channel_profile: str = 'exponential',
                 channel_truncation: str = 'trunc',
                 is_complex: bool = True,
                 device: str = 'cuda') -> None:
        # same checks, same messages' meaning as config.py:40-44
        assert channel_profile in _PROFILES, "channel_profile has to be 'exponential' or 'uniform'"
        assert channel_truncation in _TRUNCATIONS, "channel_truncation has to be 'trunc', 'tail' or 'cyclic'"
        assert channel_length > 0, "channel_length needs to be at least 1"
        assert generator_mode in _MODES, "generator_mode needs to be 'segmented' or 'random' or 'sparc'"
        assert alphabet in _ALPHABETS, "alphabet has to be one of " + ",".join(_ALPHABETS)

        self.device = device

        # dimensions (config.py:49-52)
        self.B, self.Lin = batch, block_length
        self.Nt, self.Na, self.Nr = N_transmit_antenna, N_active_antenna, N_receive_antenna
        self.sparsity = self.Na / self.Nt
        self.mode = generator_mode

        # channel (config.py:55-68)
        self.is_complex = is_complex
        self.Lh = channel_length
        self.profile = channel_profile
        self.trunc = channel_truncation
        self.Lout = self.Lin + self.Lh - 1 if channel_truncation == 'tail' else self.Lin
        self.ISI = self.Lh > 1

        # message statistics (config.py:71-76); QUIRK: Ps is split over the K symbols of a modulated alphabet
        self.Ns = self.B * self.Lin * self.Na
        self.N0 = self.B * self.Lin * (self.Nt - self.Na)
        self.alphabet = alphabet
        self.modulated = alphabet != 'OOK'
        points, gray = _constellation(alphabet)
        self.gray = gray
        self.Ps = self.sparsity / len(points) if self.modulated else self.sparsity
        self.P0 = 1 - self.sparsity
        if alphabet in _FORCES_COMPLEX:
            self.is_complex = True

        # unit average power over the listed points (config.py:117)
        self.symbols = np.array(points) / np.sqrt(np.mean(np.abs(points) ** 2))
        self.K = len(self.symbols)
        self.symbol_bits = int(np.log2(self.K))

        # rates (config.py:121-144).  QUIRK: 'segmented'/'random' count the symbol bits once per time slot,
        # 'sparc' counts Na*log2(M*K); this changes code_rate and therefore the SNR <-> Eb/N0 mapping.
        if self.mode == 'random':
            self.index_bits = np.log2(np.prod([1 + (self.Nt - self.Na) / j for j in range(1, self.Na + 1)]))
            self.info_bits = self.symbol_bits + self.index_bits
            self.code_rate = self.Lin * self.info_bits / self.Nr / self.Lout
        else:
            assert self.Nt % self.Na == 0, 'Na must divide Nt'
            if self.mode == 'segmented':
                self.index_bits = self.Na * np.log2(self.Nt / self.Na)
                self.info_bits = self.symbol_bits + self.index_bits
                self.code_rate = self.Lin * self.info_bits / self.Nr / self.Lout
            else:  # 'sparc'
                self.Mc, self.Mr = self.Nt, self.Nr
                self.Lc, self.Lr = self.Lin, self.Lout
                self.index_bits = self.Na * np.log2(self.Nt // self.Na)
                self.inner_code_rate = self.Na * np.log2((self.Nt // self.Na) * self.K) / self.Mr
                self.code_rate = self.Lc * self.inner_code_rate / self.Lr
        # section geometry: the reference defines M, L, n only for 'sparc'; defined for every sectioned mode here
        # because the kernels need them (a superset of the reference attributes).
        if self.Nt % self.Na == 0:
            self.M = self.Nt // self.Na
            self.L = self.Na * self.Lin
        self.n = self.Nr * self.Lout
        self.N = self.Nt * self.Lin

        # iteration budget and limits (config.py:147-154)
        self.N_Layers = iterations
        self.kappa = self.Lout / self.Lin
        with np.errstate(divide='ignore', invalid='ignore'):
            self.min_amp_snr = 1 / (self.kappa * (1 / (np.exp(2 * self.code_rate) - 1) - 1 / self.Lh))
        self.min_snr = 2 ** self.code_rate - 1
        self.min_snr_dB = 10 * np.log10(self.min_snr)
        self.shannon_limit_dB = self.min_snr_dB - 10 * np.log10(self.code_rate)

        # results directory key (config.py:157)
        self.name = (f'{self.alphabet},{self.mode}/{self.profile},{self.trunc}/'
                     f'Nt={self.Nt},Na={self.Na},Nr={self.Nr},Lh={self.Lh},Lin={self.Lin}')

    # convenience used by the detectors; not part of the reference surface
    @property
    def loss_index_bits(self) -> int:
        """Bits kept by Loss.de2bi for the index XOR (loss.py:20)."""
        return int(math.ceil(math.log2(self.Lin * self.B * self.Na))) if self.Lin * self.B * self.Na > 0 else 0
