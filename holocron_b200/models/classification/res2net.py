"""Res2Net on the fused kernels — API mirror of holocron/models/classification/res2net.py (ScaleConv2d :22-76,
Bottle2neck :79-135, _res2net :138-157, res2net50_26w_4s :179-205).

A Bottle2neck is a ResNet bottleneck whose 3x3 unit is replaced by ``scale - 1`` chained 3x3 units on ``width``-channel
slices of the tensor (26 channels at stage 1: the convolution kernels zero-pad them to 32 internally), the previous slice's
output being added to the next slice's input; the last slice is passed through (average-pooled when the block strides).
Module tree, parameter names and init order are the reference's; every ``conv -> BN -> ReLU`` runs as a fused unit, the
1x1 expansion's BatchNorm pass also adds the shortcut and applies the block's final activation."""
import math
from typing import Any, Callable, List, Optional

import torch
from torch import Tensor, nn

from .._blocks import FusedSequential
from ..utils import _configure_model, _requested_checkpoint, conv_sequence
from .resnet import ResNet, _ResBlock

__all__ = ["Bottle2neck", "ScaleConv2d", "res2net50_26w_4s"]


class ScaleConv2d(nn.Module):
    """Hierarchical 3x3 convolutions over channel slices (reference res2net.py:22-76)."""

    def __init__(self, scale: int, planes: int, kernel_size: int, stride: int = 1, groups: int = 1, downsample: bool = False,
                 act_layer: Optional[nn.Module] = None, norm_layer: Optional[Callable[[int], nn.Module]] = None,
                 drop_layer: Optional[Callable[..., nn.Module]] = None) -> None:
        super().__init__()
        self.scale = scale
        self.width = planes // scale
        self.conv = nn.ModuleList([
            FusedSequential(*conv_sequence(self.width, self.width, act_layer, norm_layer, drop_layer, kernel_size=3, stride=stride,
                                           padding=1, groups=groups, bias=(norm_layer is None)))
            for _ in range(max(1, scale - 1))
        ])
        self.downsample = nn.AvgPool2d(kernel_size=3, stride=stride, padding=1) if downsample else None

    def forward(self, x: Tensor) -> Tensor:
        split_x = torch.split(x, self.width, 1)
        out: List[Tensor] = []
        for idx, layer in enumerate(self.conv):
            # a down-sampling block does not chain the slices (their resolutions differ)
            res = split_x[idx] if idx == 0 or self.downsample is not None else out[-1] + split_x[idx]
            out.append(layer(res))
        if self.scale > 1:
            out.append(split_x[-1] if self.downsample is None else self.downsample(split_x[-1]))
        return torch.cat(out, 1)


class Bottle2neck(_ResBlock):
    """1x1 reduce -> ScaleConv2d -> 1x1 expand x4 (reference res2net.py:79-135)."""

    expansion: int = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None, groups: int = 1,
                 base_width: int = 26, dilation: int = 1, act_layer: Optional[nn.Module] = None,
                 norm_layer: Optional[Callable[[int], nn.Module]] = None,
                 drop_layer: Optional[Callable[..., nn.Module]] = None, scale: int = 4) -> None:
        if norm_layer is None:
            norm_layer = nn.BatchNorm2d
        if act_layer is None:
            act_layer = nn.ReLU(inplace=True)
        downsample_ = stride > 1 or downsample is not None     # the pass-through slice then needs the average pool
        width = math.floor(planes * (base_width / 64.0)) * groups
        super().__init__(
            [*conv_sequence(inplanes, width * scale, act_layer, norm_layer, drop_layer, kernel_size=1, stride=1,
                            bias=(norm_layer is None)),
             ScaleConv2d(scale, width * scale, 3, stride, groups, downsample_, act_layer, norm_layer, drop_layer),
             *conv_sequence(width * scale, planes * self.expansion, None, norm_layer, drop_layer, kernel_size=1, stride=1,
                            bias=(norm_layer is None))],
            downsample, act_layer)


def _res2net(pretrained: bool, checkpoint: Any, num_blocks: List[int], out_chans: List[int], width_per_group: int, scale: int,
             **kwargs: Any) -> ResNet:
    checkpoint = _requested_checkpoint(pretrained, checkpoint)
    model = ResNet(Bottle2neck, num_blocks, out_chans, width_per_group=width_per_group,  # type: ignore[arg-type]
                   block_args={"scale": scale}, **kwargs)
    return _configure_model(model, checkpoint)


def res2net50_26w_4s(pretrained: bool = False, checkpoint: Any = None, progress: bool = True, **kwargs: Any) -> ResNet:
    """Res2Net-50 26w x 4s (https://arxiv.org/abs/1904.01169) — reference res2net.py:179-205."""
    return _res2net(pretrained, checkpoint, [3, 4, 6, 3], [64, 128, 256, 512], 26, 4, **kwargs)
