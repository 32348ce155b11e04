"""Drop-in for the hot-path helpers of RT/utilities.py: warp, flow_warp_mask(threshold),
gram_matrix (divides by H*W only - SURVEY.md Q4), vgg_normalize (out of place)."""
from __future__ import annotations

import torch

from .. import ops


def warp(x, flo, padding_mode="zeros"):
    """RT/utilities.py:59-77 (identical to RC's)."""
    if padding_mode != "zeros":
        raise NotImplementedError("only padding_mode='zeros' is used by the reference's hot path")
    return ops.warp(x, flo)


def flow_warp_mask(flo01, flo10, padding_mode="zeros", threshold=2):
    """RT/utilities.py:80-110."""
    if padding_mode != "zeros":
        raise NotImplementedError("only padding_mode='zeros' is used by the reference's hot path")
    return ops.flow_warp_mask(flo01, flo10, float(threshold))


def gram_matrix(y: torch.Tensor):
    """F F^T / (H*W) (RT/utilities.py:155-160)."""
    b, ch, h, w = y.shape
    return ops.gram(y, 1.0 / (h * w))


def vgg_normalize(batch: torch.Tensor):
    """RT/utilities.py:163-169 - does not modify its argument."""
    return ops.vgg_normalize(batch, inplace_div=False)
