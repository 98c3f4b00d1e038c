"""PyConvResNet on the fused kernels — API mirror of holocron/models/classification/pyconv_resnet.py (PyBottleneck :34-92,
PyHGBottleneck :95-96, _pyconvresnet :99-125, factories :128-181).

A ResNet bottleneck whose 3x3 unit is a pyramidal convolution (:class:`holocron_b200.nn.PyConv2d`: parallel 3x3 / 5x5 / 7x7
/ 9x9 grouped convolutions, concatenated); no max-pool behind the stem. 1x1 units, BatchNorm + activation (+ shortcut) passes
and the dense pyramid level run on the fused kernels, the grouped levels are library calls."""
from typing import Any, Callable, List, Optional, Type, Union

from torch.nn import Module

from ...nn import PyConv2d
from ..utils import conv_sequence
from .resnet import ResNet, _ResBlock

__all__ = ["PyBottleneck", "PyHGBottleneck", "pyconv_resnet50", "pyconvhg_resnet50"]


class PyBottleneck(_ResBlock):
    """reference pyconv_resnet.py:34-92."""

    expansion: int = 4

    def __init__(self, inplanes: int, planes: int, stride: int = 1, downsample: Optional[Module] = None,
                 groups: Optional[List[int]] = None, base_width: int = 64, dilation: int = 1,
                 act_layer: Optional[Module] = None, norm_layer: Optional[Callable[[int], Module]] = None,
                 drop_layer: Optional[Callable[..., Module]] = None, num_levels: int = 2, **kwargs: Any) -> None:
        if groups is None:
            groups = [1]
        width = int(planes * (base_width / 64.0)) * min(groups)
        no_norm = norm_layer is None
        super().__init__(
            [*conv_sequence(inplanes, width, act_layer, norm_layer, drop_layer, kernel_size=1, stride=1, bias=no_norm, **kwargs),
             *conv_sequence(width, width, act_layer, norm_layer, drop_layer, conv_layer=PyConv2d, kernel_size=3, stride=stride,
                            padding=dilation, groups=groups, bias=no_norm, dilation=dilation, num_levels=num_levels, **kwargs),
             *conv_sequence(width, planes * self.expansion, None, norm_layer, drop_layer, kernel_size=1, stride=1, bias=no_norm,
                            **kwargs)],
            downsample, act_layer)


class PyHGBottleneck(PyBottleneck):
    expansion: int = 2


def _pyconvresnet(pretrained: bool, block: Type[Union[PyBottleneck, PyHGBottleneck]], num_blocks: List[int],
                  out_chans: List[int], width_per_group: int, groups: List[List[int]], **kwargs: Any) -> ResNet:
    if pretrained:
        raise NotImplementedError("the released checkpoints need network access; load a reference state_dict instead "
                                  "(the module tree and parameter names are identical)")
    model = ResNet(block, num_blocks, out_chans, stem_pool=False, width_per_group=width_per_group,  # type: ignore[arg-type]
                   block_args=[{"num_levels": len(group), "groups": group} for group in groups], **kwargs)
    model.default_cfg = None
    return model


def pyconv_resnet50(pretrained: bool = False, progress: bool = True, **kwargs: Any) -> ResNet:
    """PyConvResNet-50 (https://arxiv.org/abs/2006.11538) — reference pyconv_resnet.py:128-152."""
    return _pyconvresnet(pretrained, PyBottleneck, [3, 4, 6, 3], [64, 128, 256, 512], 64,
                         [[1, 4, 8, 16], [1, 4, 8], [1, 4], [1]], **kwargs)


def pyconvhg_resnet50(pretrained: bool = False, progress: bool = True, **kwargs: Any) -> ResNet:
    """PyConvHGResNet-50 — reference pyconv_resnet.py:155-181."""
    return _pyconvresnet(pretrained, PyHGBottleneck, [3, 4, 6, 3], [128, 256, 512, 1024], 2,
                         [[32, 32, 32, 32], [32, 64, 64], [32, 64], [32]], **kwargs)
