"""DropBlock2d module — mirrors holocron/nn/modules/dropblock.py:14-41."""
from torch import Tensor, nn

from .. import functional as F

__all__ = ["DropBlock2d"]


class DropBlock2d(nn.Module):
    """DropBlock (https://arxiv.org/abs/1810.12890). As in the reference the module hands ``p / block_size**2`` to the
    functional, which divides by ``block_size**2`` again (effective seed probability ``p / block_size**4``)."""

    def __init__(self, p: float = 0.1, block_size: int = 7, inplace: bool = False) -> None:
        super().__init__()
        self.p = p
        self.block_size = block_size
        self.inplace = inplace

    @property
    def drop_prob(self) -> float:
        return self.p / self.block_size**2

    def forward(self, x: Tensor) -> Tensor:
        return F.dropblock2d(x, self.drop_prob, self.block_size, self.inplace, self.training)

    def extra_repr(self) -> str:
        return f"p={self.p}, block_size={self.block_size}, inplace={self.inplace}"
