"""Loss modules on the hot path — mirrors holocron/nn/modules/loss.py (_Loss :25-47, FocalLoss :50-84, DiceLoss :195-219,
PolyLoss :222-246)."""
from typing import Any, List, Optional, Union

import torch
from torch import Tensor, nn

from .. import functional as F

__all__ = ["DiceLoss", "FocalLoss", "PolyLoss"]


class _Loss(nn.Module):
    """Base class: registers the class weights as a ``weight`` buffer (float w -> ``[w, 1 - w]``, list, or tensor)."""

    def __init__(self, weight: Optional[Union[float, List[float], Tensor]] = None, ignore_index: int = -100,
                 reduction: str = "mean") -> None:
        super().__init__()
        self.weight: Optional[Tensor]
        if isinstance(weight, (float, int)):
            self.register_buffer("weight", torch.Tensor([weight, 1 - weight]))
        elif isinstance(weight, list):
            self.register_buffer("weight", torch.Tensor(weight))
        elif isinstance(weight, Tensor):
            self.register_buffer("weight", weight)
        else:
            self.weight = None
        self.ignore_index = ignore_index
        if reduction not in ["none", "mean", "sum"]:
            raise NotImplementedError("argument reduction received an incorrect input")
        self.reduction = reduction


class FocalLoss(_Loss):
    """Focal loss (https://arxiv.org/abs/1708.02002): ``-(1 - p_t)^gamma * w_t log p_t`` on the fused kernel."""

    def __init__(self, gamma: float = 2.0, **kwargs: Any) -> None:
        super().__init__(**kwargs)
        self.gamma = gamma

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        return F.focal_loss(x, target, self.weight, self.ignore_index, self.reduction, self.gamma)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(gamma={self.gamma}, reduction='{self.reduction}')"


class DiceLoss(_Loss):
    """Dice loss (https://arxiv.org/abs/1606.04797) on probabilities. As in the reference only ``weight`` reaches the
    base class, so ``reduction`` always reads ``'mean'``."""

    def __init__(self, weight: Optional[Union[float, List[float], Tensor]] = None, gamma: float = 1.0,
                 eps: float = 1e-8) -> None:
        super().__init__(weight)
        self.gamma = gamma
        self.eps = eps

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        return F.dice_loss(x, target, self.weight, self.gamma, self.eps)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(reduction='{self.reduction}', gamma={self.gamma}, eps={self.eps})"


class PolyLoss(_Loss):
    """Poly-1 loss (https://arxiv.org/abs/2204.12511): ``-log p_t + eps (1 - p_t)``, hard or soft targets."""

    def __init__(self, *args: Any, eps: float = 2.0, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        self.eps = eps

    def forward(self, x: Tensor, target: Tensor) -> Tensor:
        return F.poly_loss(x, target, self.eps, self.weight, self.ignore_index, self.reduction)

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}(eps={self.eps}, reduction='{self.reduction}')"
