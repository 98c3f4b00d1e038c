"""Batch collation on the input side of the training step — API mirror of holocron/utils/data/collate.py (Mixup :16-64).

The reference's training scripts wrap ``default_collate`` with ``Mixup`` (references/classification/train.py:133-136,
``--mixup-alpha 0.2`` by default): it runs in the DataLoader workers on HOST tensors and hands the step a mixed image batch plus
soft (N, K) targets. Same contract here - host tensors in, host tensors out, and the same random draws in the same order (one
Beta sample, then one permutation), so a seeded data pipeline yields bit-identical batches; the soft targets feed
``torch.nn.CrossEntropyLoss`` / :class:`holocron_b200.nn.PolyLoss` unchanged. Only tensor ops: device tensors work as well."""
from typing import Tuple

import torch
from torch import Tensor, nn
from torch.distributions.beta import Beta

__all__ = ["Mixup"]


class Mixup(nn.Module):
    """MixUp (https://arxiv.org/abs/1710.09412) as a collate function: ``mix(*default_collate(batch))``.

    Args:
        num_classes: number of classes (1: binary targets become an (N, 1) column)
        alpha: parameter of the Beta(alpha, alpha) mixing law; 0 switches the mixing off (targets are still encoded)
    """

    def __init__(self, num_classes: int, alpha: float = 0.2) -> None:
        if alpha < 0:
            raise ValueError("`alpha` only takes positive values")
        super().__init__()
        self.num_classes, self.alpha = num_classes, alpha

    def _soft_targets(self, targets: Tensor, dtype: torch.dtype) -> Tensor:
        """(N,) class indices -> (N, K) one-hot rows, or an (N, 1) column for a single class; 2-D targets pass through."""
        if targets.ndim == 1 and self.num_classes > 1:
            targets = nn.functional.one_hot(targets, num_classes=self.num_classes)
        elif targets.ndim == 1 and self.num_classes == 1:
            targets = targets.unsqueeze(1)
        return targets.to(dtype=dtype)

    @staticmethod
    def _blend_(batch: Tensor, order: Tensor, lam: Tensor) -> Tensor:
        """In place: batch <- lam * batch + (1 - lam) * batch[order] (the gathered partner is scaled first, like the reference)."""
        partner = batch[order]
        partner.mul_(1 - lam)
        return batch.mul_(lam).add_(partner)

    def forward(self, inputs: Tensor, targets: Tensor) -> Tuple[Tensor, Tensor]:
        targets = self._soft_targets(targets, inputs.dtype)
        if self.alpha == 0:
            return inputs, targets
        lam = Beta(self.alpha, self.alpha).sample()            # draw 1
        order = torch.randperm(inputs.size()[0])               # draw 2
        return self._blend_(inputs, order, lam), self._blend_(targets, order, lam)
