"""Activation modules mirroring holocron/nn/modules/activation.py."""
from typing import ClassVar, List

import torch
from torch import Tensor, nn

from .. import functional as F

__all__ = ["FReLU", "HardMish", "NLReLU"]


class _Activation(nn.Module):
    __constants__: ClassVar[List[str]] = ["inplace"]

    def __init__(self, inplace: bool = False) -> None:
        super().__init__()
        self.inplace = inplace

    def extra_repr(self) -> str:
        return "inplace=True" if self.inplace else ""


class HardMish(_Activation):
    """f(x) = x/2 * min(2, max(0, x + 2)) — reference activation.py:28-38."""

    def forward(self, x: Tensor) -> Tensor:
        return F.hard_mish(x, inplace=self.inplace)


class NLReLU(_Activation):
    """f(x) = ln(1 + max(0, x)) (beta fixed to 1 as in the reference module, activation.py:41-55)."""

    def forward(self, x: Tensor) -> Tensor:
        return F.nl_relu(x, inplace=self.inplace)


class FReLU(nn.Module):
    """Funnel activation max(x, BN(dwconv_kxk(x))) — reference activation.py:58-82.

    Children ``conv`` (depth-wise, with bias) and ``bn`` keep the reference's parameter names/shapes; the forward
    runs the fused depth-wise-conv kernels (see ``holocron_b200.nn._dwconv``).
    """

    def __init__(self, in_channels: int, kernel_size: int = 3) -> None:
        super().__init__()
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size, padding=kernel_size // 2, groups=in_channels)
        self.bn = nn.BatchNorm2d(in_channels)

    def forward(self, x: Tensor) -> Tensor:
        from .._dwconv import frelu_forward
        return frelu_forward(x, self.conv, self.bn)
