"""B200-native latent hot path of the PIC/REM progressive codec.

Public surface = the reference's own API for this path (drop-in):
    ChannelMask, ste_round                      (layers/channel_mask.py)
    EntropyModel, GaussianConditional, EntropyBottleneck   (entropy_models/entropy_models.py)
    get_scale_table                             (models/pic.py:12-17)
plus the fused per-slice operator ``progressive_slice_forward`` and the spatially tiled
multi-GPU select (``distributed``), and the host-side coder after the path (``codec``: CDF tables,
rANS streams, progressive level packing; include/pic_codec.h).  All compute runs in libpic_latent.so (hand-written
sm_100a CUDA behind the C ABI of include/pic_latent.h); importing this package never builds
or falls back: a missing library raises at first use.
"""
import math

import torch

from . import _lib, codec, distributed, ops
from ._lib import LIB_PATH, build, lib
from .channel_mask import ChannelMask, ste_round
from .entropy_models import EntropyBottleneck, EntropyModel, GaussianConditional, LowerBound
from .functional import lrp_merge, progressive_slice_forward, rate_bpp, rem_merge

SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    """models/pic.py:16-17."""
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


__all__ = ["ChannelMask", "ste_round", "EntropyModel", "EntropyBottleneck", "GaussianConditional", "LowerBound",
           "progressive_slice_forward", "rate_bpp", "lrp_merge", "rem_merge", "get_scale_table", "ops", "codec", "distributed", "build", "lib", "LIB_PATH"]
