"""Activation layers of the hot path: ``HardMish``, ``NLReLU`` (single fused pointwise pass each, csrc/pointwise.cu) and
``FReLU`` (depth-wise conv + BatchNorm + funnel max on the fused kernels).

Same constructors, attribute names, child modules and ``repr`` as holocron/nn/modules/activation.py:28-82.
"""
from typing import Callable, Optional

from torch import Tensor, nn

from .. import functional as F

__all__ = ["FReLU", "HardMish", "NLReLU"]


class _Pointwise(nn.Module):
    """Stateless element-wise layer: subclasses only name the functional they apply."""

    _fn: Optional[Callable[..., Tensor]] = None
    __constants__ = ["inplace"]

    def __init__(self, inplace: bool = False) -> None:
        super().__init__()
        self.inplace = bool(inplace)

    def forward(self, x: Tensor) -> Tensor:
        return type(self)._fn(x, inplace=self.inplace)

    def extra_repr(self) -> str:
        return "inplace=True" * self.inplace


class HardMish(_Pointwise):
    """``x / 2 * clamp(x + 2, 0, 2)`` (reference activation.py:28-38)."""

    _fn = staticmethod(F.hard_mish)


class NLReLU(_Pointwise):
    """``log(1 + relu(x))``; like the reference module (activation.py:41-55) it always uses ``beta = 1``."""

    _fn = staticmethod(F.nl_relu)


class FReLU(nn.Module):
    """Funnel activation ``max(x, BN(dwconv_kxk(x)))`` (reference activation.py:58-82).

    ``conv`` (depth-wise, biased) and ``bn`` are plain torch modules so that parameter names and shapes stay those of
    the reference; the arithmetic runs in ``holocron_b200.nn._dwconv.frelu_forward``.
    """

    def __init__(self, in_channels: int, kernel_size: int = 3) -> None:
        super().__init__()
        half = kernel_size // 2
        self.conv = nn.Conv2d(in_channels, in_channels, kernel_size, padding=half, groups=in_channels)
        self.bn = nn.BatchNorm2d(in_channels)

    def forward(self, x: Tensor) -> Tensor:
        from .._dwconv import frelu_forward
        return frelu_forward(x, self.conv, self.bn)
