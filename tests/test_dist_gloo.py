"""world_size-2 gloo test of the multi-GPU plumbing: frame sharding and the counter all-reduce."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import amp_sparc_spatialmodulation_b200 as pkg
from amp_sparc_spatialmodulation_b200 import _cabi
from amp_sparc_spatialmodulation_b200.dist import allreduce_counters, shard_range
from conftest import config_from_meta, load_golden
from oracle import loss_oracle as lo


def test_shard_range_partitions_frames():
    for frames in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 4, 8):
            spans = [shard_range(frames, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == frames
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = load_golden("bamp_c1")
    F = g["x"].shape[0]
    cfg = config_from_meta(g["meta"], batch=F)
    lo_f, hi_f = shard_range(F, rank, world)
    N = g["x"].shape[1]
    idx = g["idx"].reshape(F, -1) + (np.arange(F) * N)[:, None]
    # each rank scores its own frame range (the oracle stands in for the kernel on CPU), with the call-wide
    # truncation of the index bits and global flat indices, exactly as the sharded GPU path does
    c = lo.error_counters(g["xmap"][lo_f:hi_f], g["xmmse"][lo_f:hi_f], g["x"][lo_f:hi_f], g["sym"][lo_f:hi_f].ravel(),
                          idx[lo_f:hi_f].ravel() - lo_f * N, cfg.symbols, cfg.gray, dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin),
                          iters=g["iters"][lo_f:hi_f], index_bits_kept=0)
    buf = torch.zeros(_cabi.NUM_COUNTERS, dtype=torch.int64)
    for i, k in enumerate(_cabi.COUNTER_NAMES):
        buf[i] = c[k]
    buf[16:20] = torch.tensor([c[k] for k in _cabi.SQERR_NAMES], dtype=torch.float64).view(torch.int64)
    allreduce_counters(buf)
    if rank == 0:
        out.put(buf.numpy().copy())
    dist.barrier()
    dist.destroy_process_group()


def test_counter_allreduce_world2_matches_single_process():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    merged = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    g = load_golden("bamp_c1")
    F = g["x"].shape[0]
    cfg = config_from_meta(g["meta"], batch=F)
    N = g["x"].shape[1]
    gidx = (g["idx"].reshape(F, -1) + (np.arange(F) * N)[:, None]).ravel()
    whole = lo.error_counters(g["xmap"], g["xmmse"], g["x"], g["sym"].ravel(), gidx, cfg.symbols, cfg.gray,
                              dict(Nt=cfg.Nt, Na=cfg.Na, Lin=cfg.Lin), iters=g["iters"], index_bits_kept=0)
    got = _cabi.counters_to_dict(merged)
    for k in _cabi.COUNTER_NAMES:
        assert got[k] == whole[k], k
    for k in _cabi.SQERR_NAMES:
        assert abs(got[k] - whole[k]) <= 1e-9 * max(1.0, whole[k])
    L = pkg.Loss(cfg)
    L.record(got, got["iters"] / got["frames"])
    assert float(L.loss["fer"]) == whole["frame_err"] / F
