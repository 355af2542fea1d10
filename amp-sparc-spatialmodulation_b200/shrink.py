"""``Shrink`` -- element-wise denoiser family of the reference (/root/reference/shrink.py:8-166).

In the reference only ``vamp2.py`` (not wired into any driver) consumes this class, and only in ``random`` mode,
which is outside the sectioned hot path.  The constructor surface is kept so imports
written against the reference resolve; calling it raises until the ``random``-mode kernels land.
"""
from torch import nn

from ._cabi import AmpsmError
from .config import Config

_KINDS = ("bayes", "shrink", "lasso", "shrinkOOK")


class Shrink(nn.Module):
    def __init__(self, config: Config, shrink_fn: str) -> None:
        super().__init__()
        assert shrink_fn in _KINDS, "shrink_fn needs to be one of " + ", ".join(_KINDS)
        self.config, self.kind = config, shrink_fn

    def forward(self, r, cov):
        raise AmpsmError(f"Shrink('{self.kind}') has no sm_100a kernel yet: 'random'-mode denoisers are a later row "
                         "of the hot-path scope (SURVEY.md section 8f); there is no CPU fallback")
