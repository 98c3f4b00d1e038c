"""Batch collation on the input side of the training step — API mirror of holocron/utils/data/collate.py (Mixup :16-64).

The reference's training scripts wrap ``default_collate`` with ``Mixup`` (references/classification/train.py:133-136,
``--mixup-alpha 0.2`` by default): it runs in the DataLoader workers on HOST tensors and hands the step a mixed image batch
plus soft (N, K) targets. Same contract here (host tensors in, host tensors out, same RNG draws in the same order: one Beta
sample, one permutation), so that a seeded data pipeline produces the same batches; the soft targets feed
``torch.nn.CrossEntropyLoss`` / :class:`holocron_b200.nn.PolyLoss` unchanged. Works on device tensors too (all tensor ops)."""
from typing import Tuple

import torch
from torch import Tensor
from torch.distributions.beta import Beta
from torch.nn.functional import one_hot

__all__ = ["Mixup"]


class Mixup(torch.nn.Module):
    """MixUp (https://arxiv.org/abs/1710.09412) as a collate function: ``mix(*default_collate(batch))``.

    Args:
        num_classes: number of classes (1: binary targets become an (N, 1) column)
        alpha: parameter of the Beta(alpha, alpha) mixing distribution; 0 disables mixing (targets are still one-hot encoded)
    """

    def __init__(self, num_classes: int, alpha: float = 0.2) -> None:
        super().__init__()
        self.num_classes = num_classes
        if alpha < 0:
            raise ValueError("`alpha` only takes positive values")
        self.alpha = alpha

    def forward(self, inputs: Tensor, targets: Tensor) -> Tuple[Tensor, Tensor]:
        if targets.ndim == 1:                     # class indices -> (N, K) one-hot rows / (N, 1) column
            if self.num_classes > 1:
                targets = one_hot(targets, num_classes=self.num_classes)
            elif self.num_classes == 1:
                targets = targets.unsqueeze(1)
        targets = targets.to(dtype=inputs.dtype)
        if self.alpha == 0:
            return inputs, targets
        lam = Beta(self.alpha, self.alpha).sample()
        index = torch.randperm(inputs.size()[0])
        # x <- lam * x + (1 - lam) * x[perm], same for the targets; in place on the batch like the reference
        partner_x, partner_t = inputs[index, :], targets[index]
        inputs.mul_(lam).add_(partner_x.mul_(1 - lam))
        targets.mul_(lam).add_(partner_t.mul_(1 - lam))
        return inputs, targets
