"""B200-native BAMP / SCAMP / VAMP spatial-modulation detectors behind the reference's Python API.

``Config / Channel / Data / Loss / BAMP / SCAMP / VAMP`` keep the call signatures of
AhmedKishki/AMP-SPARC-SpatialModulation; the detectors run hand-written sm_100a CUDA kernels through the C-ABI
library ``csrc/libampsm_b200.so`` (see include/ampsm_b200.h).  There is no CPU fallback: the detectors raise
when the library or a CUDA device is missing.
"""
from .config import Config
from .channel import Channel
from .data import Data
from .loss import Loss
from .bamp import BAMP
from .scamp import SCAMP
from .vamp import VAMP, svd_batched
from .vamp2 import VAMP as VAMP2
from .shrink import Shrink
from .framegen import FrameStream
from .simulate import MonteCarlo, device_frames, run_scamp

__all__ = ["Config", "Channel", "Data", "Loss", "BAMP", "SCAMP", "VAMP", "VAMP2", "Shrink", "svd_batched", "FrameStream", "MonteCarlo", "device_frames",
           "run_scamp"]
