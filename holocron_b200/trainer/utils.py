"""Model-freezing helpers of the training loop — host-side mirrors of holocron/trainer/utils.py (freeze_bn :14-30,
freeze_model :33-71, split_normalization_params :74-101): same names, arguments and effects."""
from typing import List, Optional, Tuple

from torch import nn
from torch.nn.modules.batchnorm import _BatchNorm

__all__ = ["freeze_bn", "freeze_model", "split_normalization_params"]


def freeze_bn(mod: nn.Module) -> None:
    """BatchNorm layers whose affine parameters are all frozen stop updating their statistics and normalise with the
    running ones (reference utils.py:14-30). The fused kernels honour this per layer: a frozen BatchNorm inside a training
    model takes the running-statistics path (scale / shift folded once, no statistics reduction)."""
    for m in mod.modules():
        if isinstance(m, _BatchNorm) and m.affine and all(not p.requires_grad for p in m.parameters()):
            m.track_running_stats = False
            m.eval()


def freeze_model(model: nn.Module, last_frozen_layer: Optional[str] = None, frozen_bn_stat_update: bool = False) -> None:
    """Freezes every parameter up to and including ``last_frozen_layer`` (parameters are assumed to be registered in
    forward order), unfreezes the rest (reference utils.py:33-71)."""
    for p in model.parameters():
        p.requires_grad_(True)
    if isinstance(last_frozen_layer, str):
        reached = False
        for n, p in model.named_parameters():
            if not reached or n.startswith(last_frozen_layer):
                p.requires_grad_(False)
            if n.startswith(last_frozen_layer):
                reached = True
            elif reached:
                break
        if not reached:
            raise ValueError(f"Unable to locate child module {last_frozen_layer}")
    if not frozen_bn_stat_update:
        freeze_bn(model)


def split_normalization_params(model: nn.Module, norm_classes: Optional[List[type]] = None
                               ) -> Tuple[List[nn.Parameter], List[nn.Parameter]]:
    """(normalisation-layer parameters, all other parameters), trainable ones only (reference utils.py:74-101)."""
    if not norm_classes:
        norm_classes = [_BatchNorm, nn.LayerNorm, nn.GroupNorm]
    for t in norm_classes:
        if not issubclass(t, nn.Module):
            raise ValueError(f"Class {t} is not a subclass of nn.Module.")
    classes = tuple(norm_classes)
    norm_params: List[nn.Parameter] = []
    other_params: List[nn.Parameter] = []
    for module in model.modules():
        if next(module.children(), None):
            other_params.extend(p for p in module.parameters(recurse=False) if p.requires_grad)
        elif isinstance(module, classes):
            norm_params.extend(p for p in module.parameters() if p.requires_grad)
        else:
            other_params.extend(p for p in module.parameters() if p.requires_grad)
    return norm_params, other_params
