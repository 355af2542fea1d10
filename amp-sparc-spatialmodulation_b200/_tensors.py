"""Tensor plumbing between torch and the C-ABI: the pointers handed to the kernels must address plain, dense memory."""
import torch


def dense(t: torch.Tensor, dev, dtype, *shape) -> torch.Tensor:
    """Tensor as the kernels read it: on `dev`, of `dtype`, optionally reshaped, contiguous, and with torch's lazy
    conjugate / negative bits MATERIALISED -- `torch.linalg.svd` on CUDA returns Vh as a conj-view (`V.mH`), whose
    `data_ptr()` addresses the un-conjugated memory; `.contiguous()` alone keeps the bit."""
    t = t.to(dev, dtype).resolve_conj().resolve_neg()
    if shape:
        t = t.reshape(*shape)
    return t.contiguous()


def ptr(t):
    return None if t is None else t.data_ptr()
