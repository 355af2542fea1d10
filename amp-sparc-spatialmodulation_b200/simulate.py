"""Monte-Carlo SNR sweep -- the GPU counterpart of ``Model.simulate`` in the reference's drivers
(/root/reference/bamp_model.py:44-67, vamp_model.py:46-69, scamp_model.py:45-66; the reference's own simulate.py is empty).

The reference loops ``for EbN0: for epoch: [new channel every res] -> message -> y = A x + noise -> detector ->
Loss.accumulate`` with ``batch=1``.  Here one SNR point is a pool of independent frames, generated ON THE DEVICE
(channel, message, noise: torch's Philox generators, seeded per point and rank), detected in chunks by the sm_100a kernels
and scored by their fused Loss epilogue; frames are sharded over the ranks of ``torch.distributed`` by contiguous ranges
(one process per GPU) and the ONLY exchange is one all-reduce of the 24-word counter block per SNR point
(SURVEY.md section 8e).  Rank 0 turns the counters into the reference's 14 rates and writes ``<path>/<EbN0dB>.json`` with
the reference's schema (loss.py:304-323); the sweep stops early once FER < 1e-3 (bamp_model.py:66).

Channels (one draw per frame):
  'iid'        H ~ CN(0, 1/Nr) i.i.d.; with Lin > 1 or Lh > 1 the frame is an ISI block: Lh tap matrices
               ~ CN(0, pdp_l Lout / (Nr Lin)) as channel.py:53-55, applied as a (truncated / tail / cyclic) block
               convolution -- the dense block-Toeplitz matrix is never formed (BAMP.detect_taps)
  'kronecker'  H = Rr^(1/2) G Rt^(1/2), G i.i.d. CN(0, 1/Nr), exponential correlation R[i,j] = rho^|i-j|
               -- BASELINE.json config 5; the reference has no correlated generator
Detectors: 'bamp' (H as is), 'vamp' (batched Jacobi SVD of every H on the device + iterations, one C-ABI call).
SCAMP shares one design matrix per call and keeps the reference's own ``Channel.generate_as_sparc`` (see ``run_scamp``).
"""
import argparse
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist

from . import _cabi
from .bamp import BAMP
from .channel import Channel
from .config import Config
from .data import Data
from .dist import allreduce_counters, shard_range
from .loss import Loss
from .scamp import SCAMP
from .vamp import VAMP


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def exp_corr_root(m: int, rho: float, device) -> torch.Tensor:
    """Hermitian square root of the exponential correlation matrix R[i,j] = rho^|i-j| (complex64, on `device`)."""
    i = torch.arange(m, dtype=torch.float64)
    R = torch.as_tensor(rho, dtype=torch.float64) ** (i[:, None] - i[None, :]).abs()
    w, V = torch.linalg.eigh(R)
    return ((V * w.clamp_min(0).sqrt()) @ V.T).to(device=device, dtype=torch.complex64)


def device_frames(cfg: Config, frames: int, snr: float, gen: torch.Generator, channel='iid', rho_t=0.0, rho_r=0.0):
    """One chunk of frames on the generator's device: per-frame H (frames, n, N), y (frames, n), x (frames, N), Gray labels
    (frames * L,) and flat non-zero positions (frames * L,) as Data.generate_message returns them (data.py:74-91)."""
    if cfg.Lin != 1 or cfg.Lh != 1:
        if channel != 'iid':
            raise _cabi.AmpsmError("ISI frames (Lin > 1 or Lh > 1) are generated with i.i.d. taps only")
        return device_isi_frames(cfg, frames, snr, gen)
    dev = gen.device
    n, N, M, L = cfg.n, cfg.N, cfg.M, cfg.L
    H = torch.view_as_complex(torch.randn(frames, n, N, 2, device=dev, generator=gen) * float(np.sqrt(1 / cfg.Nr / 2)))
    if channel == 'kronecker':
        H = exp_corr_root(n, rho_r, dev) @ H @ exp_corr_root(N, rho_t, dev)
    elif channel != 'iid':
        raise ValueError(f"unknown channel model {channel!r}")
    ant = torch.randint(0, M, (frames, L), device=dev, generator=gen)
    k = torch.randint(0, cfg.K, (frames, L), device=dev, generator=gen)
    sym = torch.as_tensor(np.asarray(cfg.symbols)).to(dev, torch.complex64)
    gray = torch.as_tensor(np.asarray(cfg.gray)).to(dev, torch.int64)
    pos = ant + torch.arange(L, device=dev) * M
    x = torch.zeros(frames, N, dtype=torch.complex64, device=dev)
    x.scatter_(1, pos, sym[k])
    sigma2 = (cfg.Na / cfg.Nr) / snr                                       # bamp.py:111,134
    noise = torch.view_as_complex(torch.randn(frames, n, 2, device=dev, generator=gen) * float(np.sqrt(sigma2 / 2)))
    cols = torch.gather(H, 2, pos[:, None, :].expand(frames, n, L))        # the L active columns of every frame
    y = (cols * sym[k][:, None, :]).sum(-1) + noise
    idx = (pos + torch.arange(frames, device=dev)[:, None] * N).reshape(-1).contiguous()
    return H.contiguous(), y.contiguous(), x, gray[k].reshape(-1).contiguous(), idx


def device_isi_frames(cfg: Config, frames: int, snr: float, gen: torch.Generator):
    """ISI frames on the generator's device: per-frame taps (frames, Lh, Nr, Nt) with the scaling of channel.py:55
    (uniform or exponential power-delay profile), y = block convolution of the message with the taps ('trunc', 'tail' or
    'cyclic', channel.py:56-72) + noise.  Returns (taps, y, x, labels, flat indices) like device_frames; the first element
    goes to BAMP.detect_taps instead of BAMP.detect."""
    dev = gen.device
    Nt, Nr, Lin, Lout, Lh, M, L, N, n = cfg.Nt, cfg.Nr, cfg.Lin, cfg.Lout, cfg.Lh, cfg.M, cfg.L, cfg.N, cfg.n
    pdp = np.exp(-np.arange(Lh)) if cfg.profile == 'exponential' else np.ones(Lh)
    pdp = pdp / pdp.sum()
    scale = torch.as_tensor(np.sqrt(pdp * Lout / Nr / Lin / 2), dtype=torch.float32, device=dev)
    taps = torch.view_as_complex((torch.randn(frames, Lh, Nr, Nt, 2, device=dev, generator=gen)
                                  * scale[None, :, None, None, None]).contiguous())
    ant = torch.randint(0, M, (frames, L), device=dev, generator=gen)
    k = torch.randint(0, cfg.K, (frames, L), device=dev, generator=gen)
    sym = torch.as_tensor(np.asarray(cfg.symbols)).to(dev, torch.complex64)
    gray = torch.as_tensor(np.asarray(cfg.gray)).to(dev, torch.int64)
    pos = ant + torch.arange(L, device=dev) * M                            # section s covers entries [s M, (s+1) M)
    x = torch.zeros(frames, N, dtype=torch.complex64, device=dev)
    x.scatter_(1, pos, sym[k])
    xs = x.view(frames, Lin, Nt)
    y = torch.zeros(frames, Lin + Lh - 1, Nr, dtype=torch.complex64, device=dev)     # full linear convolution
    for l in range(Lh):                                                    # out slot i = in slot j + l
        y[:, l:l + Lin] += torch.einsum('frt,fjt->fjr', taps[:, l], xs)
    if cfg.trunc == 'cyclic':                                              # the post-transient wraps to the first slots
        y[:, :Lh - 1] += y[:, Lin:]
    y = y[:, :Lout].reshape(frames, n)
    sigma2 = (cfg.Na / cfg.Nr) / snr
    y = y + torch.view_as_complex(torch.randn(frames, n, 2, device=dev, generator=gen) * float(np.sqrt(sigma2 / 2)))
    idx = (pos + torch.arange(frames, device=dev)[:, None] * N).reshape(-1).contiguous()
    return taps, y.contiguous(), x, gray[k].reshape(-1).contiguous(), idx


class MonteCarlo:
    """SNR sweep of one detector over device-generated frames, sharded over the ranks of torch.distributed."""

    def __init__(self, config: Config, algorithm='bamp', frames_per_point=1 << 20, chunk=1 << 18, channel='iid', rho_t=0.0,
                 rho_r=0.0, seed=1234, path=None, device=None, generator='torch', **detector_kw):
        self.config, self.algorithm = config, algorithm
        self.generator = generator
        self.frames_per_point, self.chunk = int(frames_per_point), int(chunk)
        self.channel, self.rho_t, self.rho_r, self.seed = channel, rho_t, rho_r, seed
        self.path = path
        self.rank, self.world = _world()
        self.device = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.rate = config.code_rate
        self.min_snr = config.shannon_limit_dB                              # bamp_model.py:27
        if (config.Lin != 1 or config.Lh != 1) and algorithm != 'bamp':
            raise ValueError("ISI frames (Lin > 1 or Lh > 1) run through BAMP's structured operator only")
        if algorithm == 'bamp':
            self.amp = BAMP(config, outputs=False, **detector_kw)
        elif algorithm == 'vamp':
            self.amp = VAMP(config, outputs=False, **detector_kw)
        else:
            raise ValueError("MonteCarlo runs 'bamp' or 'vamp'; SCAMP shares a design matrix per call: see run_scamp()")
        self.loss = Loss(config)
        # generator='kernel': the library's own Philox stream (framegen.py) -- one generation kernel for BAMP; for VAMP the frames
        # are drawn inside the Jacobi SVD kernel and the channel matrix never exists in HBM
        self.stream = None
        if generator == 'kernel':
            from .framegen import FrameStream
            self.stream = FrameStream(config, seed=seed, channel=channel, rho_t=rho_t, rho_r=rho_r, device=self.device)
        elif generator != 'torch':
            raise ValueError("generator is 'torch' (torch's Philox generators, several passes) or 'kernel' (csrc/framegen.cuh)")

    def run_point(self, EbN0dB: float, point_index: int = 0) -> dict:
        """All frames of one SNR point: returns the GLOBAL counter dict (summed over ranks)."""
        SNRdB = EbN0dB + 10 * np.log10(self.rate)
        snr = 10 ** (SNRdB / 10)
        lo, hi = shard_range(self.frames_per_point, self.rank, self.world)
        gen = torch.Generator(device=self.device).manual_seed(self.seed + 7919 * point_index + 104729 * self.rank)
        total = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=self.device)
        f0 = lo
        while f0 < hi:
            nf = min(self.chunk, hi - f0)
            if self.stream is not None:
                first = point_index * self.frames_per_point + f0               # global frame number in the stream
                if self.algorithm == 'vamp':
                    det = self.amp.detect_generated(self.stream, first, nf, snr)
                else:
                    H, y, x, lab, idx = self.stream.frames(first, nf, snr)
                    det = self.amp.detect(H, y, snr, x, lab, idx, frame_base=0)
                total[:16] += det.counters[:16]
                total[16:20] = (total[16:20].view(torch.float64) + det.counters[16:20].view(torch.float64)).view(torch.int64)
                f0 += nf
                continue
            H, y, x, lab, idx = device_frames(self.config, nf, snr, gen, self.channel, self.rho_t, self.rho_r)
            if self.algorithm == 'bamp' and H.dim() == 4:                  # ISI frames: H holds the taps
                det = self.amp.detect_taps(H, y, snr, x, lab, idx, cyclic=self.config.trunc == 'cyclic', frame_base=0)
            elif self.algorithm == 'bamp':
                det = self.amp.detect(H, y, snr, x, lab, idx, frame_base=0)
            else:
                det = self.amp.detect_from_channel(H, y, snr, x, lab, idx, frame_base=0)
            total[:16] += det.counters[:16]
            total[16:20] = (total[16:20].view(torch.float64) + det.counters[16:20].view(torch.float64)).view(torch.int64)
            f0 += nf
        allreduce_counters(total)                                          # the one collective of the SNR point
        return _cabi.counters_to_dict(total.cpu().numpy())

    def simulate(self, final=None, start=None, step: float = 1.0, stop_fer: float = 1e-3):
        """bamp_model.py:44-67: sweep Eb/N0 from `start` (default: ceil of the Shannon limit) to `final` (default
        start + 20 dB); one JSON per point on rank 0; stop once FER < stop_fer.  Returns the list of per-point dicts."""
        if start is None:
            start = int(np.ceil(self.min_snr))
        if final is None:
            final = start + 20.0
        out = []
        for pi, EbN0dB in enumerate(np.arange(start, final + step, step)):
            SNRdB = EbN0dB + 10 * np.log10(self.rate)
            torch.cuda.synchronize(self.device)
            t0 = time.perf_counter()
            c = self.run_point(float(EbN0dB), pi)              # ends with the all-reduce and a host read: synchronised
            seconds = time.perf_counter() - t0
            self.loss.dump()
            self.loss.loss = {}
            self.loss.record(c, c['iters'] / max(c['frames'], 1))
            rates = {k: float(np.asarray(self.loss.loss[k])) for k in self.loss.keys}
            point = dict(EbN0dB=float(EbN0dB), SNRdB=float(SNRdB), T=c['iters'] / max(c['frames'], 1), frames=c['frames'],
                         nan_frames=c['nan_frames'], seconds=seconds, frames_per_s=c['frames'] / seconds, ranks=self.world, **rates)
            out.append(point)
            if self.rank == 0:
                print(f"EbN0dB={EbN0dB:g} frames={c['frames']} FER={rates['fer']:.3e} ier={rates['ier']:.3e} "
                      f"ber={rates['ber']:.3e} iter={point['T']:.2f} {seconds:.2f} s ({c['frames'] / seconds:.3e} frames/s on {self.world} rank(s), "
                      f"generation + SVD + detection)", flush=True)
                if self.path:
                    os.makedirs(self.path, exist_ok=True)
                    self.loss.export(SNRdB, float(EbN0dB), self.path)
            if rates['fer'] < stop_fer:
                break
        return out


def run_scamp(config: Config, epochs: int, EbN0dB: float, res: int = 1, seed: int = 0, **detector_kw) -> dict:
    """scamp_model.py:45-66 for one SNR point: `epochs` frames in groups of `res` that share one design matrix drawn by
    the reference's own generator (channel.py:75-95); each group is ONE SCAMP call with batch = res."""
    np.random.seed(seed)
    torch.manual_seed(seed)
    SNRdB = EbN0dB + 10 * np.log10(config.code_rate)
    snr = 10 ** (SNRdB / 10)
    total = {}
    done = 0
    while done < epochs:
        nf = min(res, epochs - done)
        cfg = Config(config.Nt, config.Na, config.Nr, config.Lin, config.Lh, batch=nf, generator_mode=config.mode,
                     iterations=config.N_Layers, alphabet=config.alphabet, channel_profile=config.profile,
                     channel_truncation=config.trunc, device=config.device)
        ch, da = Channel(cfg), Data(cfg)
        W, A = ch.generate_as_sparc()
        x, sym, idx = da.generate_message()
        y = A @ x + ch.awgn(snr)
        det = SCAMP(cfg, outputs=False, **detector_kw).detect(W, A, y, snr, x, sym, idx)
        c = det.counters_dict()
        for k, v in c.items():
            total[k] = total.get(k, 0) + v
        done += nf
    return total


def main(argv=None):
    ap = argparse.ArgumentParser(description="Monte-Carlo Eb/N0 sweep on the GPU (one process per GPU under torchrun)")
    ap.add_argument("--alg", default="bamp", choices=["bamp", "vamp"])
    ap.add_argument("--Nt", type=int, default=64)
    ap.add_argument("--Na", type=int, default=1)
    ap.add_argument("--Nr", type=int, default=32)
    ap.add_argument("--Lin", type=int, default=1, help="time slots per frame (block length)")
    ap.add_argument("--Lh", type=int, default=1, help="channel taps (ISI when > 1)")
    ap.add_argument("--truncation", default="trunc", choices=["trunc", "tail", "cyclic"])
    ap.add_argument("--alphabet", default="16QAM")
    ap.add_argument("--iterations", type=int, default=20)
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per SNR point (all ranks together)")
    ap.add_argument("--generator", default="torch", choices=["torch", "kernel"],
                    help="kernel: the library's Philox stream (VAMP: frames drawn inside the SVD kernel, H never in HBM)")
    ap.add_argument("--chunk", type=int, default=1 << 18)
    ap.add_argument("--channel", default="iid", choices=["iid", "kronecker"])
    ap.add_argument("--rho-t", type=float, default=0.0)
    ap.add_argument("--rho-r", type=float, default=0.0)
    ap.add_argument("--ebn0-start", dest="start", type=float, default=None, help="first Eb/N0 in dB (default: ceil of the Shannon limit)")
    ap.add_argument("--ebn0-final", dest="final", type=float, default=None, help="last Eb/N0 in dB (default: first + 20)")
    ap.add_argument("--ebn0-step", dest="step", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--path", default=None, help="directory for the per-point JSON files (reference schema)")
    a = ap.parse_args(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cfg = Config(a.Nt, a.Na, a.Nr, a.Lin, a.Lh, batch=a.chunk, generator_mode='sparc', iterations=a.iterations,
                 alphabet=a.alphabet, channel_profile='uniform', channel_truncation=a.truncation, device=f"cuda:{local}")
    mc = MonteCarlo(cfg, a.alg, frames_per_point=a.frames, chunk=a.chunk, channel=a.channel, rho_t=a.rho_t, rho_r=a.rho_r,
                    seed=a.seed, path=a.path, generator=a.generator)
    pts = mc.simulate(final=a.final, start=a.start, step=a.step)
    if mc.rank == 0:
        print(json.dumps(pts))
        if a.path:
            with open(os.path.join(a.path, "sweep_record.json"), "w") as f:
                json.dump(dict(args=vars(a), world_size=world, points=pts), f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
