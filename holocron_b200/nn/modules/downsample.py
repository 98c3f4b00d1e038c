"""Pooling modules on the hot path — mirrors holocron/nn/modules/downsample.py (ConcatDownsample2d :26-40, GlobalAvgPool2d :58-77, SPP :154-167)."""
from typing import List

import torch
from torch import Tensor, nn

__all__ = ["ConcatDownsample2d", "GlobalAvgPool2d", "SPP"]


class ConcatDownsample2d(nn.Module):
    """Stacks adjacent pixels on the channel axis (reference downsample.py:26-40; YOLOv2's pass-through layer)."""

    def __init__(self, scale_factor: int) -> None:
        super().__init__()
        self.scale_factor = scale_factor

    def forward(self, x: Tensor) -> Tensor:
        from ..functional import concat_downsample2d
        return concat_downsample2d(x, self.scale_factor)


class GlobalAvgPool2d(nn.Module):
    """Global average pooling over the spatial dims (RepVGG / ReXNet / Darknet heads, SE blocks).

    bf16 channels_last activations (what the fused conv blocks produce) go through the NHWC pooling kernel with
    fp32 accumulation; any other CUDA tensor goes through the generic row-reduction kernel.
    """

    def __init__(self, flatten: bool = False) -> None:
        super().__init__()
        self.flatten = flatten

    def forward(self, x: Tensor) -> Tensor:
        from .._fused import global_avg_pool_flat
        out = global_avg_pool_flat(x)
        if self.flatten:
            return out
        return out.view(x.size(0), x.size(1), 1, 1)

    def extra_repr(self) -> str:
        return "flatten=True" if self.flatten else ""


class SPP(nn.ModuleList):
    """Spatial pyramid pooling: cat(x, maxpool_k(x) for k in kernel_sizes) along channels (YOLOv4 neck)."""

    def __init__(self, kernel_sizes: List[int]) -> None:
        super().__init__([nn.MaxPool2d(k_size, stride=1, padding=k_size // 2) for k_size in kernel_sizes])

    def forward(self, x: Tensor) -> Tensor:
        feats = [x] + [pool_layer(x) for pool_layer in self]
        return torch.cat(feats, dim=1)
