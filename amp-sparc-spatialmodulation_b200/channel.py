"""Channel-side input generators for the detectors (inputs to the hot path, not part of it).

Same public surface and the same random-number call order as the reference ``Channel``
(/root/reference/channel.py:8-116): with identical ``np.random.seed`` / ``torch.manual_seed`` the matrices and the
noise produced here are bit-identical to the reference's, which is what the golden fixtures under
``tests/golden`` rely on.  ``draw_frames`` is an addition: it stacks per-frame draws for the batched kernels.
"""
import numpy as np
import torch

from .config import Config


class Channel:
    def __init__(self, config: Config) -> None:
        self.device = config.device
        self.B, self.Lin = config.B, config.Lin
        self.Nt, self.Na, self.Nr = config.Nt, config.Na, config.Nr
        self.Lh, self.Lout = config.Lh, config.Lout
        self.trunc = config.trunc
        self.sparsity = config.sparsity
        self.is_complex = config.is_complex
        # power-delay profile (channel.py:27-31); QUIRK: profile 'random' passes Config's assert but leaves
        # the reference without a pdp attribute -- here it fails at construction with a clear message instead.
        if config.profile == 'exponential':
            pdp = np.exp(-np.arange(self.Lh))
        elif config.profile == 'uniform':
            pdp = np.ones(self.Lh)
        else:
            raise AttributeError("channel_profile 'random' has no power-delay profile (reference channel.py:27-31)")
        self.pdp = pdp / pdp.sum()
        self.dtype = torch.complex64 if self.is_complex else torch.float32
        self.npdtype = np.complex64 if self.is_complex else np.float32

    # -- taps ---------------------------------------------------------------------------------------------
    def _taps(self):
        """Two successive normal draws, real part first (channel.py:53-54 / 85-86)."""
        shape = (self.Nr, self.Nt, self.Lh)
        re = np.random.normal(size=shape)
        im = np.random.normal(size=shape)
        return re + 1j * im

    def _wrap_rows(self, H):
        """Row blocks that hold the channel's post-transient response (channel.py:60-72)."""
        Nr, Nt, Lh = self.Nr, self.Nt, self.Lh
        tail = H[-Nr:, -Nt * Lh:-Nt]
        return [tail[:, :Nt * (Lh - l - 1)] for l in range(Lh - 1)]

    def _as_taps(self, h):
        """(Nr, Nt, Lh) numpy taps -> (Lh, Nr, Nt) tensor, the layout of ampsm_bamp_detect_taps / BAMP.detect_taps."""
        return torch.tensor(np.ascontiguousarray(np.moveaxis(h, -1, 0)), dtype=self.dtype, requires_grad=False,
                            device=self.device)

    def generate_channel(self, return_taps: bool = False) -> torch.Tensor:
        """Block-Toeplitz MIMO-ISI matrix, (Nr*Lout) x (Nt*Lin) (channel.py:40-73).  ``return_taps=True`` (an addition)
        also returns the scaled taps (Lh, Nr, Nt) the matrix is built from -- same random draws either way."""
        Nr, Nt, Lh, Lin = self.Nr, self.Nt, self.Lh, self.Lin
        h = self._taps() * np.sqrt(self.pdp * self.Lout / Nr / Lin / 2)
        H = np.zeros((Lin * Nr, Lin * Nt), dtype=self.npdtype)
        for l in range(Lh):
            H += np.kron(np.eye(Lin, Lin, -l), h[:, :, l])
        if self.trunc == 'tail':
            extra = np.zeros((Nr * (Lh - 1), Nt * Lin), dtype=self.npdtype)
            for l, blk in enumerate(self._wrap_rows(H)):
                extra[l * Nr:(l + 1) * Nr, -blk.shape[1]:] = blk
            H = np.block([[H], [extra]])
        elif self.trunc == 'cyclic':
            for l, blk in enumerate(self._wrap_rows(H)):
                H[l * Nr:(l + 1) * Nr, -blk.shape[1]:] = blk
        H = torch.tensor(H, dtype=self.dtype, requires_grad=False, device=self.device)
        return (H, self._as_taps(h)) if return_taps else H

    def generate_as_sparc(self, return_taps: bool = False):
        """Base matrix W (Lout x Lin) and design matrix A with per-block variance W (channel.py:75-95).
        ``return_taps=True`` (an addition) appends the taps sqrt(W[l, 0]) h_l, (Lh, Nr, Nt), A is built from."""
        W = np.zeros((self.Lout, self.Lin))
        for l in range(self.Lh):
            W += np.eye(self.Lout, self.Lin, -l) * self.pdp[l]
        W = W / W.mean() * self.Na / self.Nr
        h = self._taps() / np.sqrt(2 * self.Na * self.Lin)
        A = np.zeros((self.Nr * self.Lout, self.Nt * self.Lin), dtype=self.npdtype)
        for l in range(self.Lh):
            A += np.kron(np.eye(self.Lout, self.Lin, -l) * np.sqrt(W), h[:, :, l])
        taps = h * np.sqrt(np.array([W[l, 0] if l < self.Lout else 0.0 for l in range(self.Lh)]))
        W = torch.tensor(W, dtype=torch.float32, requires_grad=False, device=self.device)
        A = torch.tensor(A, dtype=self.dtype, requires_grad=False, device=self.device)
        return (W, A, self._as_taps(taps)) if return_taps else (W, A)

    def generate_as_random(self) -> torch.Tensor:
        """i.i.d. CN(0, 1/(Lin*Nr)) matrix from the torch generator (channel.py:97-101)."""
        shape = (self.Nr * self.Lout, self.Nt * self.Lin)
        re = torch.normal(mean=0, std=1, size=shape, device=self.device)
        im = torch.normal(mean=0, std=1, size=shape, device=self.device)
        return ((re + 1j * im) / np.sqrt(2 * self.Lin * self.Nr)).to(self.dtype)

    def awgn(self, SNR) -> torch.Tensor:
        """Noise of variance (Na/Nr)/SNR per receive sample, shape (B, Nr*Lout, 1) (channel.py:103-116)."""
        shape = (self.B, self.Nr * self.Lout, 1)
        re = torch.normal(mean=0., std=1., size=shape, device=self.device)
        im = torch.normal(mean=0., std=1., size=shape, device=self.device)
        return (re + 1j * im) * np.sqrt(self.Na / self.Nr / SNR / 2)
