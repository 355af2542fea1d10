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
    """Sum counter blocks over ranks in place with ONE collective.  Words 0..15 are integers, 16..19 float64 sums (bit-cast
    in the int64 tensor): the block travels as 24 float64 words -- integer counts are exact in float64 up to 2^53 (9e15; a
    1e8-frame sweep point counts at most 1e8 x Lin x Na x bits events) -- and is converted back on arrival."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters
    assert counters.dtype == torch.int64 and counters.numel() == _cabi.NUM_COUNTERS
    packed = counters.to(torch.float64)
    packed[16:20] = counters[16:20].view(torch.float64)
    dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    out = packed.round().to(torch.int64)
    out[16:20] = packed[16:20].contiguous().view(torch.int64)
    counters.copy_(out)
    return counters


def merge_counter_dicts(dicts):
    """Host-side sum of counter dicts (e.g. chunks of one SNR point)."""
    out = {}
    for d in dicts:
        for k, v in d.items():
            out[k] = out.get(k, 0) + v
    return out
