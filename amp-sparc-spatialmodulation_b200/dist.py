"""Multi-GPU plumbing: frames are independent, so the path shards by contiguous frame ranges with no data-path
collective; the only exchange is one all-reduce (SUM) of the 24-word counter block per SNR point
(SURVEY.md section 8e).  Works with the ``nccl`` backend on GPUs and ``gloo`` on CPU (tests)."""
import numpy as np
import torch
import torch.distributed as dist

from . import _cabi


def shard_range(frames: int, rank: int, world: int):
    """Contiguous frame range [lo, hi) of `rank`; the first `frames % world` ranks hold one extra frame."""
    base, extra = divmod(int(frames), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def allreduce_counters(counters: torch.Tensor, group=None) -> torch.Tensor:
    """Sum counter blocks over ranks in place.  Words 0..15 are integers, 16..19 float64 sums (bit-cast in the
    int64 tensor), so the two halves are reduced with their own dtypes."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters
    assert counters.dtype == torch.int64 and counters.numel() == _cabi.NUM_COUNTERS
    ints = counters[:16].clone()
    sq = counters[16:20].view(torch.float64).clone()
    dist.all_reduce(ints, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(sq, op=dist.ReduceOp.SUM, group=group)
    counters[:16] = ints
    counters[16:20] = sq.view(torch.int64)
    return counters


def merge_counter_dicts(dicts):
    """Host-side sum of counter dicts (e.g. chunks of one SNR point)."""
    out = {}
    for d in dicts:
        for k, v in d.items():
            out[k] = out.get(k, 0) + v
    return out
