#!/usr/bin/env python
"""Headline benchmark: BAMP frame-iterations/s at Nt x Nr = 64 x 32, 16-QAM spatial modulation (BASELINE.json
configs[1]: 1M frames per SNR point, iterations = 20, complex64), one B200 or N of them.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step = one SNR point: every frame of the per-GPU pool is detected once (per-frame channel matrix, early exit as
the reference, fused hard decision + error counters), followed by the NCCL all-reduce of the counter block.
`value` = frame-iterations executed by all ranks / device time (CUDA events, max over ranks), inputs resident
in HBM.  `e2e` = the same metric through the C-ABI host entry point (ampsm_bamp_detect_host) with pinned host
buffers, host<->device copies inside the timed region.  `cpu_baseline` / `--impl reference` time the reference's OWN torch
CPU path (bamp.py:116-143 with batch=1, one call per frame, Loss included) on all host cores when the unmodified reference is
importable -- /root/reference in the build container, baseline/_ref (git-ignored, placed by scripts/install_reference.py)
on the GPU box -- and the numpy oracle port beside it (kind "reference" / "port"); the port alone when it is not.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NT, NA, NR, LIN, LH, ALPHABET, ITERS = 64, 1, 32, 1, 1, '16QAM', 20
FLOP_PER_FRAME_ITER = 20 * NR * NT + 18 * NT * 16 + 30 * NR + 20 * NT      # 61 632
BYTES_PER_FRAME = 8 * NR * NT + 8 * NR + 8 * NT                             # H, y, x_true: 17 152


_RESULT_FD = None


def quiet_stdout():
    """The driver reads ONE JSON line from stdout: everything else that lands there (NCCL prints its version banner to
    stdout under torchrun, libraries may print too) is sent to stderr; emit() writes the result line to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    if _RESULT_FD is None:
        print(line, flush=True)
    else:
        os.write(_RESULT_FD, (line + "\n").encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1 << 20, help="frames per GPU per step (1M = one SNR point)")
    ap.add_argument("--snr-db", type=float, default=15.0)
    ap.add_argument("--e2e-frames", type=int, default=1 << 17)
    ap.add_argument("--cpu-frames", type=int, default=0, help="frames of the CPU baseline sample (0 = auto)")
    ap.add_argument("--kernel", default="auto", choices=["auto", "generic", "fast"])
    ap.add_argument("--vamp-frames", type=int, default=1 << 18, help="frames per GPU of the VAMP leg (0 = skip it)")
    ap.add_argument("--c3-frames", type=int, default=1 << 14, help="frames per GPU of the config-3 VAMP leg (128 x 64; 0 = skip it)")
    ap.add_argument("--c3-snr-db", type=float, default=2.0)
    ap.add_argument("--scamp-frames", type=int, default=1024, help="frames of the config-4 SCAMP leg (A 1088 x 16384 shared; 0 = skip it)")
    ap.add_argument("--scamp-ebn0-db", type=float, default=6.0)
    ap.add_argument("--c1-frames", type=int, default=1 << 20, help="frames per GPU of the config-1 BAMP leg (8 x 4 QPSK; 0 = skip it)")
    ap.add_argument("--c128-frames", type=int, default=1 << 12, help="frames of the config-3 complex128 VAMP leg (0 = skip it)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fixed-t", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU side (oracle)
def _cpu_inputs(frames, snr_db, seed):
    """Synthetic frames of the C2 shape, drawn like channel.py:53-55 / data.py:74-91 / channel.py:113-115."""
    from amp_sparc_spatialmodulation_b200.config import Config
    cfg = Config(NT, NA, NR, LIN, LH, batch=frames, generator_mode='sparc', iterations=ITERS, alphabet=ALPHABET,
                 channel_profile='uniform', device='cpu')
    rng = np.random.default_rng(seed)
    n, N = cfg.n, cfg.N
    H = ((rng.standard_normal((frames, n, N), dtype=np.float32) + 1j * rng.standard_normal((frames, n, N), dtype=np.float32))
         * np.float32(np.sqrt(1 / NR / 2))).astype(np.complex64)
    ant = rng.integers(0, N, frames)
    k = rng.integers(0, cfg.K, frames)
    x = np.zeros((frames, N), np.complex64)
    x[np.arange(frames), ant] = cfg.symbols[k]
    sigma2 = (NA / NR) / 10 ** (snr_db / 10)
    noise = ((rng.standard_normal((frames, n), dtype=np.float32) + 1j * rng.standard_normal((frames, n), dtype=np.float32))
             * np.float32(np.sqrt(sigma2 / 2))).astype(np.complex64)
    y = (np.matmul(H, x[..., None])[..., 0] + noise).astype(np.complex64)
    return cfg, H, y, x, sigma2


def _cpu_worker(args):
    frames, snr_db, seed = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import amp_oracle, loss_oracle
    cfg, H, y, x, sigma2 = _cpu_inputs(frames, snr_db, seed)
    t0 = time.perf_counter()
    r = amp_oracle.bamp_detect(H, y, sigma2, cfg.symbols, cfg.L, cfg.M, ITERS, shift='section')
    loss_oracle.map_decision(r["xmap"], cfg.symbols, cfg.gray, cfg.M)
    return int(r["iters"].sum()), time.perf_counter() - t0


def reference_dir():
    """Where the unmodified reference can be imported from (None: only the oracle port is available)."""
    for d in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.exists(os.path.join(d, "bamp.py")) and os.path.exists(os.path.join(d, "loss.py")):
            return d
    return None


def _ref_worker(args):
    """The reference's torch CPU path as its drivers run it (bamp_model.py:44-67): batch=1, one BAMP.forward (= all
    iterations + Loss) per frame, per-frame channel; one torch thread per process, one process per core.  Only the
    detector call is timed (input generation excluded).  Returns (frame-iterations, detector seconds, frames)."""
    frames, snr_db, seed, refdir = args
    import torch
    torch.set_num_threads(1)
    if refdir not in sys.path:
        sys.path.insert(0, refdir)
    import config as rconfig
    import channel as rchannel
    import data as rdata
    import bamp as rbamp
    np.random.seed(seed)
    torch.manual_seed(seed)
    cfg = rconfig.Config(NT, NA, NR, LIN, LH, batch=1, generator_mode='sparc', iterations=ITERS, alphabet=ALPHABET,
                         channel_profile='uniform', device='cpu')
    ch, da, amp = rchannel.Channel(cfg), rdata.Data(cfg), rbamp.BAMP(cfg)
    snr = 10 ** (snr_db / 10)
    iters, secs = 0, 0.0
    with torch.no_grad():
        for _ in range(frames):
            H = ch.generate_channel()
            x, sym, idx = da.generate_message()
            y = H @ x + ch.awgn(snr)
            t0 = time.perf_counter()
            loss = amp(H, y, snr, x, sym, idx)
            secs += time.perf_counter() - t0
            iters += int(loss.loss['T'])
    return iters, secs, frames


class CpuPool:
    """Process pool over the host cores running the numpy oracle port; one slice of frames per process."""

    def __init__(self, workers):
        import multiprocessing as mp
        self.workers = workers
        self.pool = mp.get_context("spawn").Pool(workers)
        self.pool.map(_cpu_worker, [(8, 10.0, 1)] * workers)                # start-up + imports, untimed

    def rate(self, frames_total, snr_db, seed=7):
        per = max(1, frames_total // self.workers)
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [(per, snr_db, seed + w) for w in range(self.workers)])
        wall = time.perf_counter() - t0
        return sum(r[0] for r in res) / wall, per * self.workers, wall

    def ref_rate(self, frames_total, snr_db, refdir, seed=7):
        """All cores run the reference concurrently; the rate is the sum of the per-core detector rates."""
        per = max(1, frames_total // self.workers)
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker, [(per, snr_db, seed + w, refdir) for w in range(self.workers)])
        wall = time.perf_counter() - t0
        return sum(r[0] / r[1] for r in res), per * self.workers, wall, sum(r[0] for r in res) / max(sum(r[2] for r in res), 1)

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    """--impl reference: the reference's own torch CPU path (batch=1 per frame, all host cores) when the unmodified reference
    is importable, else the numpy oracle port; a bounded sample of the arm's workload per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    refdir = reference_dir()
    pool = CpuPool(cores)
    port_frames = 2048 * cores
    port_rate, port_used, port_wall = pool.rate(port_frames, args.snr_db)            # the port beside it, one sample
    if refdir:
        pool.ref_rate(4 * cores, args.snr_db, refdir)                               # imports + first calls, untimed
        probe, _, wall, _ = pool.ref_rate(16 * cores, args.snr_db, refdir)
        per_core = max(8, min(2048, int(16 * 2.5 / max(wall, 1e-3))))               # ~2.5 s per step
        frames = args.cpu_frames or per_core * cores
        run = lambda n, seed: pool.ref_rate(n, args.snr_db, refdir, seed=seed)[:3]
        kind = "reference"
        what = (f"the reference's own torch CPU path ({os.path.relpath(refdir, ROOT) if refdir.startswith(ROOT) else refdir}: "
                f"bamp.py BAMP.forward incl. Loss, batch=1 per frame, {cores} processes x 1 torch thread, detector time only)")
    else:
        frames = args.cpu_frames or 4096 * cores
        run = lambda n, seed: pool.rate(n, args.snr_db, seed=seed)
        kind = "port"
        what = f"numpy oracle port (oracle/amp_oracle.py), {cores} processes (the reference itself is not importable here)"
    for w in range(args.warmup):
        run(max(cores * 4, frames // 4), 50 + w)
    rates = []
    t_all = time.perf_counter()
    for s in range(args.steps):
        rate, used, wall = run(frames, 100 + s)
        rates.append(rate)
    ms = (time.perf_counter() - t_all) / max(args.steps, 1) * 1e3
    pool.close()
    v = float(np.mean(rates))
    sample = f"{used} frames/step of BAMP 64x32 16-QAM at {args.snr_db} dB through {what}"
    emit(json.dumps({
        "impl": "reference", "metric": "BAMP frame-iterations/s", "value": v, "unit": "frame-iter/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (denoiser float64)", "data": "synthetic",
        "config": workload_config(args, args.frames, max(1, args.gpus)),
        "cpu_baseline": {"value": v, "unit": "frame-iter/s", "cores": cores, "kind": kind, "sample": sample},
        "cpu_port": {"value": port_rate, "unit": "frame-iter/s", "cores": cores, "kind": "port",
                     "sample": f"{port_used} frames in {port_wall:.1f} s, numpy oracle port (vectorised over frames), {cores} processes"},
        "e2e": {"value": v, "unit": "frame-iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, frames, world):
    return {"workload": f"BAMP Nt={NT} Nr={NR} Na={NA} 16-QAM SM, per-frame i.i.d. Rayleigh H, iterations={ITERS}, "
                        f"early exit as reference, SNR {args.snr_db} dB",
            "frames_per_gpu_per_step": frames, "global_frames_per_step": frames * world, "iterations_max": ITERS,
            "snr_db": args.snr_db, "l2": "inputs (17 KB/frame) exceed L2 at >= 8k frames; no flush needed",
            "parallelism": f"frames sharded over {world} GPU(s), one NCCL all-reduce of the 24-word counter block per step"}


# ------------------------------------------------------------------------------------------------ clocks sampler
class Clocks:
    """SM clock and throttle reasons sampled DURING the timed region.  In-process NVML (nvidia_ml_py) every 2 ms -- the timed
    region of a default run is a few tens of milliseconds, shorter than one nvidia-smi polling period -- with `nvidia-smi
    -lms` as the fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index, self.nvml, self.stop = [], None, index, None, False
        self.sm, self.mx, self.reasons, self.source = [], [], set(), None
        self.t0 = self.t1 = None                       # the timed region (time.perf_counter), set by mark_start / mark_end

    def mark_start(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def __enter__(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self._physical_index()}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll_nvml(self):
        nv = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        try:
            self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM)))
        except Exception:
            pass
        while not self.stop:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                r = int(get_reasons(self.handle)) if get_reasons else 0
                self.sm.append((time.perf_counter(), mhz, frozenset(k for k, b in bits.items() if r & b)))
            except Exception:
                break
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.nvml:
            self.stop = True
            self.thread.join(timeout=2)
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        window = "timed"
        if self.nvml:
            # the sampler runs from before the warm-up: keep the samples taken inside the timed region; if the region was
            # shorter than one polling period, the samples of the warm-up steps (the same workload) stand in and say so
            inside = [x for x in self.sm if self.t0 is not None and self.t0 <= x[0] <= (self.t1 or x[0])]
            if not inside:
                inside, window = [x for x in self.sm if self.t0 is None or x[0] <= (self.t1 or x[0])][-8:], "warmup+timed"
            sm, mx = [x[1] for x in inside], self.mx
            reasons = sorted(set().union(*[x[2] for x in inside])) if inside else []
        else:
            sm = [float(r[0]) for r in self.rows if r and r[0].replace('.', '', 1).isdigit()]
            mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '', 1).isdigit()]
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "source": self.source, "window": window}


# ------------------------------------------------------------------------------------------------ GPU side
def make_gpu_inputs(torch, cfg, frames, snr_db, dev, seed):
    """Synthetic pool on the device: H ~ CN(0, 1/Nr) per frame, one active antenna with a 16-QAM symbol, AWGN."""
    gen = torch.Generator(device=dev).manual_seed(seed)
    n, N = cfg.n, cfg.N
    H = torch.empty(frames, n, N, dtype=torch.complex64, device=dev)
    Hr = torch.view_as_real(H)
    chunk = 1 << 16
    for lo in range(0, frames, chunk):
        hi = min(frames, lo + chunk)
        Hr[lo:hi].normal_(0.0, float(np.sqrt(1 / NR / 2)), generator=gen)
    ant = torch.randint(0, N, (frames,), device=dev, generator=gen)
    k = torch.randint(0, cfg.K, (frames,), device=dev, generator=gen)
    sym = torch.as_tensor(cfg.symbols).to(dev, torch.complex64)
    gray = torch.as_tensor(np.asarray(cfg.gray)).to(dev, torch.int64)
    x = torch.zeros(frames, N, dtype=torch.complex64, device=dev)
    ar = torch.arange(frames, device=dev)
    x[ar, ant] = sym[k]
    sigma2 = (NA / NR) / 10 ** (snr_db / 10)
    noise = torch.empty(frames, n, dtype=torch.complex64, device=dev)
    torch.view_as_real(noise).normal_(0.0, float(np.sqrt(sigma2 / 2)), generator=gen)
    y = H[ar, :, ant] * sym[k].unsqueeze(1) + noise
    labels = gray[k].contiguous()
    idx = (ar * N + ant).to(torch.int64).contiguous()
    return H, y.contiguous(), x, labels, idx


def main():
    args = parse()
    quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    cpu_base = cpu_port = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:       # before CUDA is touched in this process
        cores = os.cpu_count() or 1
        pool = CpuPool(cores)
        frames_cpu = args.cpu_frames or 16384 * cores      # ~5 s of host work for the port
        rate, used, wall = pool.rate(frames_cpu, args.snr_db)
        cpu_port = {"value": rate, "unit": "frame-iter/s", "cores": cores, "kind": "port",
                    "sample": f"{used} frames of the same workload in {wall:.1f} s, numpy oracle port (oracle/amp_oracle.py, vectorised "
                              f"over frames), {cores} processes"}
        refdir = reference_dir()
        if refdir:
            pool.ref_rate(4 * cores, args.snr_db, refdir)
            _, _, w0, _ = pool.ref_rate(16 * cores, args.snr_db, refdir)
            per_core = max(16, min(8192, int(16 * 15.0 / max(w0, 1e-3))))            # ~15 s of host work
            rrate, rused, rwall, rT = pool.ref_rate(per_core * cores, args.snr_db, refdir)
            cpu_base = {"value": rrate, "unit": "frame-iter/s", "cores": cores, "kind": "reference",
                        "sample": f"{rused} frames of the same workload in {rwall:.1f} s through the reference's own torch CPU path "
                                  f"(bamp.py BAMP.forward incl. Loss, batch=1 per frame, mean T = {rT:.2f}; {cores} processes x 1 torch "
                                  f"thread; detector time only, input generation excluded)"}
        else:
            cpu_base = dict(cpu_port)
        pool.close()

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    ge.build()
    import amp_sparc_spatialmodulation_b200 as pkg
    from amp_sparc_spatialmodulation_b200 import _cabi
    from amp_sparc_spatialmodulation_b200.dist import allreduce_counters

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _cabi.lib()

    frames = args.frames
    cfg = pkg.Config(NT, NA, NR, LIN, LH, batch=frames, generator_mode='sparc', iterations=ITERS, alphabet=ALPHABET,
                     channel_profile='uniform', device=str(dev))
    H, y, x, labels, idx = make_gpu_inputs(torch, cfg, frames, args.snr_db, dev, 1234 + rank)
    snr = 10 ** (args.snr_db / 10)
    amp = pkg.BAMP(cfg, kernel=args.kernel, outputs=False)
    amp_fixed = pkg.BAMP(cfg, kernel=args.kernel, outputs=False, early_exit=False)
    frame_base = rank * frames

    def step(module):
        det = module.detect(H, y, snr, x, labels, idx - 0, frame_base=0)
        return allreduce_counters(det.counters) if world > 1 else det.counters

    def timed(module, steps, warmup):
        clk = Clocks(local)
        with clk:                                       # started before the warm-up: NVML start-up stays out of the timed region
            for _ in range(warmup):
                step(module)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            lib.ampsm_launch_count(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            totals = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
            kernel_ms = []
            clk.mark_start()
            e0.record()
            for _ in range(steps):
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                det = module.detect(H, y, snr, x, labels, idx, frame_base=0)
                k1.record()
                kernel_ms.append((k0, k1))
                c = allreduce_counters(det.counters) if world > 1 else det.counters
                totals += c
            e1.record()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            clk.mark_end()
        ms = e0.elapsed_time(e1)
        launches = int(lib.ampsm_launch_count(0))
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        kms = float(np.mean([a.elapsed_time(b) for a, b in kernel_ms]))
        return ms, kms, _cabi.counters_to_dict(totals.cpu().numpy()), launches, clk.summary()

    ms, kernel_ms, c, launches, clocks = timed(amp, args.steps, args.warmup)
    # counters were all-reduced: c holds the global totals over all ranks and steps
    frame_iters = c["iters"]
    value = frame_iters / (ms * 1e-3)
    frames_done = c["frames"]
    mean_T = frame_iters / max(frames_done, 1)

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    import ctypes
    tfp = ctypes.c_double(0.0)
    lib.ampsm_probe_fp32_tflops(local, ctypes.byref(tfp))
    fp32_peak = tfp.value                                # FFMA probe kernel of the library, run in this process
    fp32_nominal = 148 * 128 * 2 * 1.965e9 / 1e12        # 148 SMs x 128 lanes x 2 flop x 1.965 GHz = 74.4
    traffic = None
    try:
        per_frame = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("bamp_c2_bytes_per_frame")
        traffic = per_frame * frames if per_frame else None       # ncu dram bytes per frame x frames of this launch
    except OSError:
        pass
    per_gpu_iters_per_launch = frame_iters / max(world, 1) / max(args.steps, 1)
    achieved_gbs = frames * BYTES_PER_FRAME / (kernel_ms * 1e-3) / 1e9
    achieved_tf = per_gpu_iters_per_launch * FLOP_PER_FRAME_ITER / (kernel_ms * 1e-3) / 1e12
    r_hbm = {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak, "peak_source": peak_src,
             "algorithmic_bytes_per_frame": BYTES_PER_FRAME}
    r_fp32 = {"achieved": achieved_tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak if fp32_peak else None,
              "peak_source": "FFMA probe kernel run in this process (ampsm_probe_fp32_tflops)", "peak_nominal": fp32_nominal,
              "frac_of_nominal": achieved_tf / fp32_nominal, "algorithmic_flop_per_frame_iter": FLOP_PER_FRAME_ITER}
    binding = "fp32" if (r_fp32["frac"] or 0.0) >= r_hbm["frac"] else "hbm"
    top = r_fp32 if binding == "fp32" else r_hbm
    roofline = {"bound": binding, "binding": binding, "kernel": "bamp_fast_kernel (one launch per step and GPU)", "achieved": top["achieved"],
                "peak": top["peak"], "unit": top["unit"], "frac": top["frac"], "hbm": r_hbm, "fp32": r_fp32, "kernel_ms": kernel_ms,
                "traffic": traffic,
                "traffic_source": "profiles/roofline_traffic.json (dram__bytes_read+write per frame from the committed ncu --set full "
                                  "capture of this kernel) x frames of one launch; a stored constant, not measured in this run"}

    out = {
        "metric": "BAMP frame-iterations/s", "value": value, "unit": "frame-iter/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "complex64 (f32 mat-vecs, compensated-f32 exponent offsets, f32 exp; f64 only in the decision fallback)", "data": "synthetic",
        "config": workload_config(args, frames, world), "mean_iterations_per_frame": mean_T,
        "frames_per_s": frames_done / (ms * 1e-3), "fer": c["frame_err"] / max(frames_done, 1),
        "ier": c["index_err"] / max(frames_done * NA * LIN, 1), "nan_frames": c["nan_frames"],
        "roofline": roofline, "gpu_launches": launches, "clocks": clocks,
    }

    if rank == 0 or world > 1:
        pass
    # fixed-T mode (exit disabled: exactly 20 iterations per frame) against the FP32 pipe
    if not args.no_fixed_t:
        ms2, kms2, c2, _, _ = timed(amp_fixed, max(2, args.steps // 2), 1)
        tf = fp32_peak
        v2 = c2["iters"] / (ms2 * 1e-3)
        ach = (frames * ITERS * FLOP_PER_FRAME_ITER) / (kms2 * 1e-3) / 1e12
        out["fixed_T"] = {"value": v2, "unit": "frame-iter/s", "iterations": ITERS, "kernel_ms": kms2,
                          "roofline_fp32": {"bound": "fp32", "achieved": ach, "peak": tf, "unit": "TFLOP/s",
                                            "frac": (ach / tf) if tf else None,
                                            "peak_source": "FFMA probe kernel run in this process (ampsm_probe_fp32_tflops)",
                                            "peak_nominal": fp32_nominal, "frac_of_nominal": ach / fp32_nominal,
                                            "algorithmic_flop_per_frame_iter": FLOP_PER_FRAME_ITER}}

    # end to end through the C-ABI host entry point: pinned host buffers, H2D/D2H inside the timed region
    fe = min(args.e2e_frames, frames)
    # pinned host buffers placed on this GPU's NUMA node (ampsm_host_alloc: allocated and first-touched by a thread bound to the
    # CPUs next to the GPU); the host entry point binds the calling thread the same way while it issues the copies
    def host_copy(tensor, dtype):
        hb = _cabi.HostBuffer(local, tuple(tensor.shape), dtype)
        torch.from_numpy(hb.array).copy_(tensor)
        return hb
    hH, hy, hx = host_copy(H[:fe], np.complex64), host_copy(y[:fe], np.complex64), host_copy(x[:fe], np.complex64)
    hl, hi = host_copy(labels[:fe], np.int64), host_copy(idx[:fe], np.int64)
    nl, na = ctypes.c_int32(0), ctypes.c_int32(0)
    numa_narrowed = int(lib.ampsm_host_numa_info(local, ctypes.byref(nl), ctypes.byref(na)))
    prob = _cabi.make_problem(cfg, fe, kernel=args.kernel)
    alpha = _cabi.make_alphabet(cfg)
    hcount = np.zeros(_cabi.NUM_COUNTERS, dtype=np.int64)

    def host_step():
        rc = lib.ampsm_bamp_detect_host(prob, alpha, fe, hH.ptr, cfg.n * cfg.N, hy.ptr, float((NA / NR) / snr), None,
                                        hx.ptr, hl.ptr, hi.ptr, None, None, None, None, None,
                                        hcount.ctypes.data, local)
        _cabi.check(rc, "ampsm_bamp_detect_host")
    for _ in range(max(1, args.warmup)):
        host_step()
    hcount[:] = 0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_step()
    t_e2e = time.perf_counter() - t0
    ce = _cabi.counters_to_dict(hcount)
    e2e_iters = torch.tensor([float(ce["iters"]), t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        it = e2e_iters[:1].clone()
        tm = e2e_iters[1:].clone()
        dist.all_reduce(it, op=dist.ReduceOp.SUM)
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        e2e_iters = torch.cat([it, tm])
    h2d = fe * (BYTES_PER_FRAME + 16)
    out["e2e"] = {"value": float(e2e_iters[0] / e2e_iters[1]), "unit": "frame-iter/s", "h2d_bytes_per_step": int(h2d),
                  "d2h_bytes_per_step": _cabi.NUM_COUNTERS * 8, "frames_per_step": fe,
                  "api": "ampsm_bamp_detect_host (C-ABI, pinned host buffers from ampsm_host_alloc, chunked copies overlapped with the kernel)",
                  "numa": {"thread_narrowed_to_gpu_local_cpus": bool(numa_narrowed), "gpu_local_cpus": int(nl.value), "allowed_cpus": int(na.value)}}
    # ---- VAMP leg: the same 64 x 32 16-QAM frames through VAMP (vamp.py:159-191) with per-frame SVD factors resident
    # in HBM (the reference's caller computes them once per channel draw, vamp_model.py:58)
    if args.vamp_frames > 0:
        fv = min(args.vamp_frames, frames)
        for hb in (hH, hy, hx, hl, hi):
            hb.free()
        Us, ss, Vs = [], [], []
        for H1 in H[:fv].split(16384):                      # thin SVD through the Hermitian eigenproblem of H H^H (float64)
            Hd = H1.to(torch.complex128)
            w, V = torch.linalg.eigh(Hd @ Hd.mH)
            w, V = w.flip(-1), V.flip(-1)
            sv = w.clamp_min(0).sqrt()
            Us.append(V.to(torch.complex64)), ss.append(sv.to(torch.float32))
            Vs.append(((V.mH @ Hd) / sv.unsqueeze(-1)).to(torch.complex64))
            del Hd, w, V, sv
        U, sv, Vh = torch.cat(Us).contiguous(), torch.cat(ss).contiguous(), torch.cat(Vs).contiguous()
        del Us, ss, Vs
        yv, xv, lv, iv = y[:fv].contiguous(), x[:fv].contiguous(), labels[:fv].contiguous(), idx[:fv].contiguous()
        cfgv = pkg.Config(NT, NA, NR, LIN, LH, batch=fv, generator_mode='sparc', iterations=ITERS, alphabet=ALPHABET,
                          channel_profile='uniform', device=str(dev))
        vout = {}
        for tag, ee in (("exit", True), ("fixed_T", False)):
            vamp = pkg.VAMP(cfgv, kernel=args.kernel if args.kernel in ("auto", "generic", "fast") else "auto", outputs=False,
                            early_exit=ee)
            for _ in range(max(1, args.warmup)):
                det = vamp.detect(U, sv, Vh, yv, snr, xv, lv, iv)
                if world > 1:
                    allreduce_counters(det.counters)      # as in the timed loop (the first collective after a pause is slow)
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64, device=dev)
            e0.record()
            for _ in range(args.steps):
                det = vamp.detect(U, sv, Vh, yv, snr, xv, lv, iv)
                tot += allreduce_counters(det.counters) if world > 1 else det.counters
            e1.record()
            torch.cuda.synchronize()
            msv = e0.elapsed_time(e1)
            if world > 1:
                tt = torch.tensor([msv], dtype=torch.float64, device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                msv = float(tt.item())
            cv = _cabi.counters_to_dict(tot.cpu().numpy())
            vout[tag] = (cv, msv)
        # from the channel matrices: batched Jacobi SVD on the device + iterations in ONE call (vamp_model.py:56-61)
        vfh = pkg.VAMP(cfgv, outputs=False)
        Hv = H[:fv]
        for _ in range(2):
            vfh.detect_from_channel(Hv, yv, snr, xv, lv, iv)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        det = vfh.detect_from_channel(Hv, yv, snr, xv, lv, iv)
        e1.record()
        torch.cuda.synchronize()
        ch = det.counters_dict()
        ms_fh = e0.elapsed_time(e1)
        # the same workload with the frames drawn inside the SVD kernel (csrc/framegen.cuh): channel, message and noise from the
        # library's Philox stream, the channel matrix never in HBM (ampsm_vamp_detect_generated)
        stream_g = pkg.FrameStream(cfgv, seed=97 + rank, device=dev)
        for _ in range(2):
            vfh.detect_generated(stream_g, 0, fv, snr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        detg = vfh.detect_generated(stream_g, 0, fv, snr)
        e1.record()
        torch.cuda.synchronize()
        cg = detg.counters_dict()
        ms_gen = e0.elapsed_time(e1)
        cv, msv = vout["exit"]
        cf, msf = vout["fixed_T"]
        vbytes = 8 * NR * NT + 8 * NR * NR + 4 * NR + 8 * NR + 8 * NT            # Vh, U, s, y, x_true: 25 472 B
        vflop = 16 * NR * NT + 18 * NT * 16 + 40 * NT + 10 * NR                  # 54 080 flop per frame-iteration
        gbs = (cv["frames"] / world) * vbytes / (msv * 1e-3) / 1e9
        try:        # ncu dram bytes per frame (profiles/) x frames of one launch
            vtraffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("vamp_c2_bytes_per_frame") * fv
        except Exception:
            vtraffic = None
        tfl = (cf["iters"] / world) * vflop / (msf * 1e-3) / 1e12
        tf32 = out.get("fixed_T", {}).get("roofline_fp32", {}).get("peak") or 0.0
        out["vamp"] = {
            "metric": "VAMP frame-iterations/s", "value": cv["iters"] / (msv * 1e-3), "unit": "frame-iter/s",
            "frames_per_gpu_per_step": fv, "frames_per_s": cv["frames"] / (msv * 1e-3),
            "mean_iterations_per_frame": cv["iters"] / max(cv["frames"], 1), "ier": cv["index_err"] / max(cv["frames"], 1),
            "nan_frames": cv["nan_frames"],
            "config": {"workload": f"VAMP Nt={NT} Nr={NR} Na={NA} 16-QAM SM, per-frame thin SVD factors (U, s, Vh) resident in HBM, "
                                   f"iterations={ITERS}, early exit as reference, SNR {args.snr_db} dB"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                         "algorithmic_bytes_per_frame": vbytes, "traffic": vtraffic},
            "fixed_T": {"value": cf["iters"] / (msf * 1e-3), "unit": "frame-iter/s", "iterations": ITERS,
                        "roofline_fp32": {"bound": "fp32", "achieved": tfl, "peak": tf32, "unit": "TFLOP/s",
                                          "frac": (tfl / tf32) if tf32 else None, "algorithmic_flop_per_frame_iter": vflop}},
            "from_channel": {"frames_per_s": world * ch["frames"] / (ms_fh * 1e-3), "value": world * ch["iters"] / (ms_fh * 1e-3),
                             "unit": "frame-iter/s", "ms": ms_fh,
                             "what": "ampsm_vamp_detect_from_h: one-sided Jacobi SVD of every frame's H (one warp per matrix) "
                                     "+ the iterations; per-rank time, not reduced over ranks"},
            "generated": {"frames_per_s": world * cg["frames"] / (ms_gen * 1e-3), "value": world * cg["iters"] / (ms_gen * 1e-3),
                          "unit": "frame-iter/s", "ms": ms_gen, "ier": cg["index_err"] / max(cg["frames"], 1),
                          "what": "ampsm_vamp_detect_generated: every frame (channel, message, noise) drawn by Philox4x32-10 inside the "
                                  "Jacobi SVD kernel, the channel matrix never in HBM, + the iterations; per-rank time"},
        }
    # ---- BASELINE config 3: VAMP Nt=128 Nr=64 Na=4 QPSK (vamp.py:159-191), per-frame SVD factors resident in HBM, through the
    # four-warps-per-frame register-resident kernel (csrc/vamp_quad.cu); per-rank device time
    if args.c3_frames > 0:
        f3 = args.c3_frames
        cfg3 = pkg.Config(128, 4, 64, 1, 1, batch=f3, generator_mode='sparc', iterations=ITERS, alphabet='QPSK',
                          channel_profile='uniform', device=str(dev))
        snr3 = 10 ** (args.c3_snr_db / 10)
        g3 = torch.Generator(device=dev).manual_seed(4321 + rank)
        from amp_sparc_spatialmodulation_b200.simulate import device_frames
        H3, y3, x3, l3, i3 = device_frames(cfg3, f3, snr3, g3)
        Us, ss, Vs = [], [], []
        for H1 in H3.split(4096):                           # thin SVD through the Hermitian eigenproblem of H H^H (float64)
            Hd = H1.to(torch.complex128)
            wv, V = torch.linalg.eigh(Hd @ Hd.mH)
            wv, V = wv.flip(-1), V.flip(-1)
            sv3 = wv.clamp_min(0).sqrt()
            Us.append(V.to(torch.complex64)), ss.append(sv3.to(torch.float32))
            Vs.append(((V.mH @ Hd) / sv3.unsqueeze(-1)).to(torch.complex64))
            del Hd, wv, V, sv3
        U3, s3, V3 = torch.cat(Us).contiguous(), torch.cat(ss).contiguous(), torch.cat(Vs).contiguous()
        del Us, ss, Vs
        flop3 = 16 * 64 * 128 + 18 * 128 * 4 + 40 * 128 + 10 * 64               # 146 048 flop per frame-iteration
        c3out = {}
        for tag, ee in (("exit", True), ("fixed_T", False)):
            v3 = pkg.VAMP(cfg3, outputs=False, early_exit=ee)
            for _ in range(2):
                v3.detect(U3, s3, V3, y3, snr3, x3, l3, i3)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                det = v3.detect(U3, s3, V3, y3, snr3, x3, l3, i3)
            e1.record()
            torch.cuda.synchronize()
            c3out[tag] = (det.counters_dict(), e0.elapsed_time(e1) / 3)
        # the same frames from their channel matrices: per-frame Jacobi SVD (one CTA per 64 x 128 matrix) + iterations in one call
        v3h = pkg.VAMP(cfg3, outputs=False)
        for _ in range(2):
            v3h.detect_from_channel(H3, y3, snr3, x3, l3, i3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        det3h = v3h.detect_from_channel(H3, y3, snr3, x3, l3, i3)
        e1.record()
        torch.cuda.synchronize()
        c3h, ms3h = det3h.counters_dict(), e0.elapsed_time(e1)
        del H3
        (ce, mse), (cf3, msf3) = c3out["exit"], c3out["fixed_T"]
        tf32 = out.get("fixed_T", {}).get("roofline_fp32", {}).get("peak") or 0.0
        out["vamp_c3"] = {
            "metric": "VAMP frame-iterations/s", "value": ce["iters"] / (mse * 1e-3), "unit": "frame-iter/s",
            "frames_per_gpu": f3, "mean_iterations_per_frame": ce["iters"] / f3, "ier": ce["index_err"] / (4 * f3),
            "nan_frames": ce["nan_frames"],
            "config": {"workload": f"VAMP Nt=128 Nr=64 Na=4 QPSK generalized SM, per-frame thin SVD factors resident in HBM, "
                                   f"iterations={ITERS}, early exit as reference, SNR {args.c3_snr_db} dB, complex64; per-rank time"},
            "roofline_fp32": {"bound": "fp32", "achieved": ce["iters"] * flop3 / (mse * 1e-3) / 1e12, "peak": tf32, "unit": "TFLOP/s",
                              "frac": (ce["iters"] * flop3 / (mse * 1e-3) / 1e12 / tf32) if tf32 else None,
                              "algorithmic_flop_per_frame_iter": flop3},
            "fixed_T": {"value": cf3["iters"] / (msf3 * 1e-3), "unit": "frame-iter/s",
                        "tflops": cf3["iters"] * flop3 / (msf3 * 1e-3) / 1e12},
            "from_channel": {"frames_per_s": c3h["frames"] / (ms3h * 1e-3), "value": c3h["iters"] / (ms3h * 1e-3), "unit": "frame-iter/s",
                             "ms": ms3h, "ier": c3h["index_err"] / (4 * f3),
                             "what": "ampsm_vamp_detect_from_h at config 3: one-sided Jacobi SVD of every frame's 64 x 128 matrix (one CTA "
                                     "per matrix) + the iterations: the per-frame decomposition of BASELINE config 3; per-rank time"},
        }
        del U3, s3, V3
    # ---- BASELINE config 1: BAMP 8 x 4 QPSK (the reference's own CPU-runnable case) through the same entry point
    if args.c1_frames > 0:
        f1 = args.c1_frames
        cfg1 = pkg.Config(8, 1, 4, 1, 1, batch=f1, generator_mode='sparc', iterations=ITERS, alphabet='QPSK',
                          channel_profile='uniform', device=str(dev))
        from amp_sparc_spatialmodulation_b200.simulate import device_frames
        snr1 = 10 ** (10.0 / 10)
        H1, y1, x1, l1, i1 = device_frames(cfg1, f1, snr1, torch.Generator(device=dev).manual_seed(77 + rank))
        amp1 = pkg.BAMP(cfg1, outputs=False)
        for _ in range(2):
            amp1.detect(H1, y1, snr1, x1, l1, i1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            det = amp1.detect(H1, y1, snr1, x1, l1, i1)
        e1.record()
        torch.cuda.synchronize()
        c1d, ms1 = det.counters_dict(), e0.elapsed_time(e1) / 3
        b1 = 8 * 4 * 8 + 8 * 4 + 8 * 8                       # H, y, x_true: 352 B per frame
        fl1 = 20 * 4 * 8 + 18 * 8 * 4 + 30 * 4 + 20 * 8      # 1 496 flop per frame-iteration
        out["bamp_c1"] = {
            "metric": "BAMP frame-iterations/s", "value": c1d["iters"] / (ms1 * 1e-3), "unit": "frame-iter/s", "frames_per_gpu": f1,
            "mean_iterations_per_frame": c1d["iters"] / f1, "fer": c1d["frame_err"] / f1, "nan_frames": c1d["nan_frames"],
            "config": {"workload": "BAMP Nt=8 Nr=4 Na=1 QPSK SM, per-frame i.i.d. Rayleigh H, iterations=20, early exit, SNR 10 dB; per-rank time"},
            "roofline": {"hbm": {"achieved": f1 * b1 / (ms1 * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                                 "frac": f1 * b1 / (ms1 * 1e-3) / 1e9 / hbm_peak, "algorithmic_bytes_per_frame": b1},
                         "fp32": {"achieved": c1d["iters"] * fl1 / (ms1 * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
                                  "frac": c1d["iters"] * fl1 / (ms1 * 1e-3) / 1e12 / fp32_peak if fp32_peak else None,
                                  "algorithmic_flop_per_frame_iter": fl1}}}
        del H1, y1, x1
    # ---- BASELINE config 4: SCAMP, L = 256 sections x M = 64, design matrix shared by the frame batch (scamp.py:77-108):
    # Config(512, 8, 32, 32, 3, 'tail', QPSK) -> A 1088 x 16384 (8.8 % non-zero: band of Lh = 3 blocks), W 34 x 32
    if args.scamp_frames > 0 and rank == 0:
        fs = args.scamp_frames
        cfg4 = pkg.Config(512, 8, 32, 32, 3, batch=fs, generator_mode='sparc', iterations=ITERS, alphabet='QPSK',
                          channel_profile='uniform', channel_truncation='tail', device=str(dev))
        np.random.seed(0)
        W4, A4 = pkg.Channel(pkg.Config(512, 8, 32, 32, 3, batch=1, generator_mode='sparc', iterations=ITERS, alphabet='QPSK',
                                        channel_profile='uniform', channel_truncation='tail', device='cpu')).generate_as_sparc()
        W4, A4 = W4.to(dev), A4.to(dev)
        g4 = torch.Generator(device=dev).manual_seed(99)
        M4, L4, N4, n4 = cfg4.M, cfg4.L, cfg4.N, cfg4.n
        ant = torch.randint(0, M4, (fs, L4), device=dev, generator=g4)
        k4 = torch.randint(0, cfg4.K, (fs, L4), device=dev, generator=g4)
        sym4 = torch.as_tensor(np.asarray(cfg4.symbols)).to(dev, torch.complex64)
        gray4 = torch.as_tensor(np.asarray(cfg4.gray)).to(dev, torch.int64)
        pos4 = ant + torch.arange(L4, device=dev) * M4
        x4 = torch.zeros(fs, N4, dtype=torch.complex64, device=dev)
        x4.scatter_(1, pos4, sym4[k4])
        snr4 = 10 ** ((args.scamp_ebn0_db + 10 * np.log10(cfg4.code_rate)) / 10)
        s24 = (cfg4.Na / cfg4.Nr) / snr4
        y4 = x4 @ A4.T + torch.view_as_complex(torch.randn(fs, n4, 2, device=dev, generator=g4) * float(np.sqrt(s24 / 2)))
        lab4 = gray4[k4].reshape(-1).contiguous()
        idx4 = (pos4 + torch.arange(fs, device=dev)[:, None] * N4).reshape(-1).contiguous()
        # TF32 tensor peak: cuBLAS TF32 GEMM 8192^3 (library call used as the probe only), best of 5
        a32 = torch.randn(8192, 8192, device=dev)
        b32 = torch.randn(8192, 8192, device=dev)
        old_tf32 = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        best = 0.0
        for rep in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a32 @ b32
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = old_tf32
        del a32, b32
        nnz = int((A4 != 0).sum())
        sout = {}
        for tag, ee in (("exit", True), ("fixed_T", False)):
            sc = pkg.SCAMP(cfg4, outputs=False, early_exit=ee)
            for _ in range(2):
                sc.detect(W4, A4, y4, snr4, x4, lab4, idx4)
            torch.cuda.synchronize()
            lib.ampsm_launch_count(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                det = sc.detect(W4, A4, y4, snr4, x4, lab4, idx4)
            e1.record()
            torch.cuda.synchronize()
            sout[tag] = (det.counters_dict(), e0.elapsed_time(e1) / 3, int(lib.ampsm_launch_count(0)) // 3)
        (cs, mss, ls), (cf4, msf4, _) = sout["exit"], sout["fixed_T"]
        flop_nz = 16 * nnz + 18 * N4 * cfg4.K                # per frame-iteration: two complex mat-vecs over the non-zero entries + denoiser
        tf_alg = cf4["iters"] * flop_nz / (msf4 * 1e-3) / 1e12
        tf_mma = cf4["iters"] * 3 * 16 * nnz / (msf4 * 1e-3) / 1e12     # 3xTF32: three tensor-core products per algorithmic one
        out["scamp_c4"] = {
            "metric": "SCAMP frame-iterations/s", "value": cs["iters"] / (mss * 1e-3), "unit": "frame-iter/s", "frames": fs,
            "ms_per_call": mss, "mean_iterations_per_frame": cs["iters"] / fs, "fer": cs["frame_err"] / fs,
            "ver": cs["slot_err"] / (fs * cfg4.Lin), "nan_frames": cs["nan_frames"], "gpu_launches_per_call": ls,
            "config": {"workload": f"SCAMP Config(512, 8, 32, 32, 3, 'tail', QPSK): L=256 sections x M=64, A {n4} x {N4} shared by "
                                   f"{fs} frames ({nnz / A4.numel():.3f} non-zero), iterations={ITERS}, early exit, Eb/N0 "
                                   f"{args.scamp_ebn0_db} dB; one GPU (rank 0)"},
            "fixed_T": {"value": cf4["iters"] / (msf4 * 1e-3), "unit": "frame-iter/s", "ms_per_call": msf4},
            "roofline_tensor": {"bound": "tensor", "achieved": tf_mma, "peak": best, "unit": "TFLOP/s (TF32)", "frac": tf_mma / best if best else None,
                                "algorithmic_tflops": tf_alg, "algorithmic_flop_per_frame_iter": flop_nz,
                                "what": "fixed-T run; achieved = 3 x 16 x nnz(A) tensor-core flop per frame-iteration (3xTF32 split of the two "
                                        "complex mat-vecs over the non-zero blocks) / time; peak = cuBLAS TF32 GEMM 8192^3 measured in this process"}}
        del W4, A4, x4, y4
    # ---- BASELINE config 3 in complex128 (the reference fed with upcast factors, vamp.py:12-28,119): FP64 pipe
    if args.c128_frames > 0 and rank == 0:
        f5 = args.c128_frames
        cfg5 = pkg.Config(128, 4, 64, 1, 1, batch=f5, generator_mode='sparc', iterations=ITERS, alphabet='QPSK',
                          channel_profile='uniform', device=str(dev))
        snr5 = 10 ** (args.c3_snr_db / 10)
        from amp_sparc_spatialmodulation_b200.simulate import device_frames
        H5, y5, x5, l5, i5 = device_frames(cfg5, f5, snr5, torch.Generator(device=dev).manual_seed(555))
        U5, s5, V5 = torch.linalg.svd(H5.to(torch.complex128), full_matrices=False)
        U5, s5, V5, y5d = U5.contiguous(), s5.contiguous(), V5.contiguous(), y5.to(torch.complex128)
        v5 = pkg.VAMP(cfg5, outputs=False)
        for _ in range(2):
            v5.detect(U5, s5, V5, y5d, snr5, x5, l5, i5)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            det = v5.detect(U5, s5, V5, y5d, snr5, x5, l5, i5)
        e1.record()
        torch.cuda.synchronize()
        c5d, ms5 = det.counters_dict(), e0.elapsed_time(e1) / 3
        flop3 = 16 * 64 * 128 + 18 * 128 * 4 + 40 * 128 + 10 * 64
        t64 = ctypes.c_double(0.0)
        if hasattr(lib, "ampsm_probe_fp64_tflops"):
            lib.ampsm_probe_fp64_tflops(local, ctypes.byref(t64))
        ach5 = c5d["iters"] * flop3 / (ms5 * 1e-3) / 1e12
        out["vamp_c3_c128"] = {
            "metric": "VAMP frame-iterations/s", "value": c5d["iters"] / (ms5 * 1e-3), "unit": "frame-iter/s", "frames": f5,
            "mean_iterations_per_frame": c5d["iters"] / f5, "nan_frames": c5d["nan_frames"],
            "config": {"workload": f"VAMP Nt=128 Nr=64 Na=4 QPSK, complex128 factors (float64 linear stage, denoiser outputs rounded to "
                                   f"complex64/float32 as vamp.py:119), SNR {args.c3_snr_db} dB; one GPU (rank 0)"},
            "roofline_fp64": {"bound": "fp64", "achieved": ach5, "peak": t64.value or None, "unit": "TFLOP/s",
                              "frac": (ach5 / t64.value) if t64.value else None, "algorithmic_flop_per_frame_iter": flop3,
                              "peak_source": "DFMA probe kernel run in this process (ampsm_probe_fp64_tflops)"}}
        del U5, s5, V5, H5
    if cpu_base:
        out["cpu_baseline"] = cpu_base
        out["cpu_port"] = cpu_port
    if rank == 0:
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
