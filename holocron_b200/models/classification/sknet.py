"""SKNet on the fused kernels — API mirror of holocron/models/classification/sknet.py (SoftAttentionLayer :39-73,
SKConv2d :76-115, SKBottleneck :118-171, factories :203-267).

A selective-kernel unit runs ``m`` grouped 3x3 paths of growing dilation (``conv -> BN -> ReLU`` each: the grouped /
dilated convolution is a library call, normalisation + activation one fused pass), squeezes their sum to (N, C, 1, 1),
and mixes the paths with a per-channel softmax over ``m`` computed by a two-layer 1x1 bottleneck. The squeeze path
works on (N, C, 1, 1) fp32 tensors with stock modules (a few kB per step); the path mix is written without the reference's
``torch.stack`` (which would copy every path into a 5-D tensor and drop the channels_last layout)."""
from typing import Any, Callable, List, Optional

import torch
from torch import Tensor, nn

from ...nn import GlobalAvgPool2d
from .._blocks import FusedSequential
from ..utils import _configure_model, _requested_checkpoint, conv_sequence
from .resnet import ResNet, _ResBlock

__all__ = ["SKBottleneck", "SKConv2d", "SoftAttentionLayer", "sknet50", "sknet101", "sknet152"]


class SoftAttentionLayer(nn.Sequential):
    """GAP -> 1x1 -> BN -> act -> 1x1(+bias) -> sigmoid (reference sknet.py:39-73)."""

    def __init__(self, channels: int, sa_ratio: int = 16, out_multiplier: int = 1, act_layer: Optional[nn.Module] = None,
                 norm_layer: Optional[Callable[[int], nn.Module]] = None,
                 drop_layer: Optional[Callable[..., nn.Module]] = None) -> None:
        super().__init__(
            GlobalAvgPool2d(flatten=False),
            *conv_sequence(channels, max(channels // sa_ratio, 32), act_layer, norm_layer, drop_layer, kernel_size=1, stride=1,
                           bias=(norm_layer is None)),
            *conv_sequence(max(channels // sa_ratio, 32), channels * out_multiplier, nn.Sigmoid(), None, drop_layer,
                           kernel_size=1, stride=1),
        )

    def forward(self, x: Tensor) -> Tensor:  # type: ignore[override]
        mods = list(self)
        y = mods[0](x).float()              # (N, C, 1, 1): the parameters of the squeeze path are fp32 masters
        for m in mods[1:]:
            y = m(y)
        return y


class SKConv2d(nn.Module):
    """Selective-kernel convolution (reference sknet.py:76-115)."""

    def __init__(self, in_channels: int, out_channels: int, m: int = 2, sa_ratio: int = 16,
                 act_layer: Optional[nn.Module] = None, norm_layer: Optional[Callable[[int], nn.Module]] = None,
                 drop_layer: Optional[Callable[..., nn.Module]] = None, **kwargs: Any) -> None:
        super().__init__()
        self.path_convs = nn.ModuleList([
            FusedSequential(*conv_sequence(in_channels, out_channels, act_layer, norm_layer, drop_layer, kernel_size=3,
                                           bias=(norm_layer is None), dilation=idx + 1, padding=idx + 1, **kwargs))
            for idx in range(m)
        ])
        self.sa = SoftAttentionLayer(out_channels, sa_ratio, m, act_layer, norm_layer, drop_layer)

    def forward(self, x: Tensor) -> Tensor:
        paths = [path_conv(x) for path_conv in self.path_convs]
        total = paths[0]
        for p in paths[1:]:
            total = total + p
        b, c = total.shape[:2]
        z = self.sa(total).view(b, len(paths), c, 1, 1)
        attention = torch.softmax(z, dim=1).to(total.dtype)
        out = attention[:, 0] * paths[0]
        for idx in range(1, len(paths)):
            out = out + attention[:, idx] * paths[idx]
        return out


class SKBottleneck(_ResBlock):
    """1x1 reduce -> SKConv2d (32 groups) -> 1x1 expand x4 (reference sknet.py:118-171)."""

    expansion: int = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[nn.Module] = None, groups: int = 32,
                 base_width: int = 64, dilation: int = 1, act_layer: Optional[nn.Module] = None,
                 norm_layer: Optional[Callable[[int], nn.Module]] = None,
                 drop_layer: Optional[Callable[..., nn.Module]] = None,
                 conv_layer: Optional[Callable[..., nn.Module]] = None, **kwargs: Any) -> None:
        width = int(planes * (base_width / 64.0)) * groups
        super().__init__(
            [*conv_sequence(inplanes, width, act_layer, norm_layer, drop_layer, conv_layer, kernel_size=1, stride=1,
                            bias=(norm_layer is None), **kwargs),
             SKConv2d(width, width, 2, 16, act_layer, norm_layer, drop_layer, groups=groups, stride=stride),
             *conv_sequence(width, planes * self.expansion, None, norm_layer, drop_layer, conv_layer, kernel_size=1, stride=1,
                            bias=(norm_layer is None), **kwargs)],
            downsample, act_layer)


def _sknet(pretrained: bool, checkpoint: Any, num_blocks: List[int], out_chans: List[int], **kwargs: Any) -> ResNet:
    checkpoint = _requested_checkpoint(pretrained, checkpoint)
    model = ResNet(SKBottleneck, num_blocks, out_chans, **kwargs)  # type: ignore[arg-type]
    return _configure_model(model, checkpoint)


def sknet50(pretrained: bool = False, checkpoint: Any = None, progress: bool = True, **kwargs: Any) -> ResNet:
    """SKNet-50 (https://arxiv.org/abs/1903.06586) — reference sknet.py:203-231."""
    return _sknet(pretrained, checkpoint, [3, 4, 6, 3], [64, 128, 256, 512], **kwargs)


def sknet101(pretrained: bool = False, checkpoint: Any = None, progress: bool = True, **kwargs: Any) -> ResNet:
    """SKNet-101 — reference sknet.py:226-242 (a checkpoint is only used together with ``pretrained``)."""
    return _sknet(pretrained, checkpoint if pretrained else None, [3, 4, 23, 3], [64, 128, 256, 512], **kwargs)


def sknet152(pretrained: bool = False, checkpoint: Any = None, progress: bool = True, **kwargs: Any) -> ResNet:
    """SKNet-152 — reference sknet.py:245-261."""
    return _sknet(pretrained, checkpoint if pretrained else None, [3, 8, 86, 3], [64, 128, 256, 512], **kwargs)
