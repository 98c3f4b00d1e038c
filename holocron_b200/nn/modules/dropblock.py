"""``DropBlock2d`` on the fused mask / apply kernels (csrc/dropblock.cu).

Public surface of holocron/nn/modules/dropblock.py:14-41: constructor ``(p=0.1, block_size=7, inplace=False)``, attributes
``p`` / ``block_size`` / ``inplace``, the ``drop_prob`` property, the ``repr`` string, identity in eval mode.
"""
from torch import Tensor, nn

from ..functional import dropblock2d

__all__ = ["DropBlock2d"]

_REPR_FIELDS = ("p", "block_size", "inplace")


class DropBlock2d(nn.Module):
    """Drops contiguous ``block_size`` x ``block_size`` regions of every feature map (Ghiasi et al., 2018).

    Reference quirk kept on purpose: the module already divides ``p`` by the block area before calling the functional,
    which divides by the block area once more - seeds are therefore drawn with probability ``p / block_size**4``.
    """

    def __init__(self, p: float = 0.1, block_size: int = 7, inplace: bool = False) -> None:
        super().__init__()
        self.p, self.block_size, self.inplace = p, block_size, inplace

    def extra_repr(self) -> str:
        return ", ".join(f"{name}={getattr(self, name)}" for name in _REPR_FIELDS)

    @property
    def drop_prob(self) -> float:
        """Public attribute of the reference (dropblock.py:33-35): ``p`` divided by the block area."""
        return self.p / self.block_size**2

    def forward(self, x: Tensor) -> Tensor:
        return dropblock2d(x, self.drop_prob, self.block_size, self.inplace, self.training)
