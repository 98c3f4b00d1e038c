"""Model-building helpers mirroring holocron/models/utils.py (conv_sequence :28-86, fuse_conv_bn :116-143)."""
import logging
from typing import Any, Callable, List, Optional, Tuple

import torch
from torch import nn

__all__ = ["conv_sequence", "fuse_conv_bn"]

logger = logging.getLogger(__name__)


def conv_sequence(
    in_channels: int,
    out_channels: int,
    act_layer: Optional[nn.Module] = None,
    norm_layer: Optional[Callable[[int], nn.Module]] = None,
    drop_layer: Optional[Callable[..., nn.Module]] = None,
    conv_layer: Optional[Callable[..., nn.Module]] = None,
    bn_channels: Optional[int] = None,
    attention_layer: Optional[Callable[[int], nn.Module]] = None,
    blurpool: bool = False,
    **kwargs: Any,
) -> List[nn.Module]:
    """Builds ``[conv(bias = norm is None), norm, act, attention, drop(inplace=True)]`` with the reference's ordering
    and bias rule. ``blurpool`` is outside the hot path and not supported here."""
    if blurpool:
        raise NotImplementedError("BlurPool2d is outside the B200 hot path (SURVEY.md §2 row 5)")
    if conv_layer is None:
        conv_layer = nn.Conv2d
    if bn_channels is None:
        bn_channels = out_channels
    # a convolution followed by a normalisation layer does not need a bias
    kwargs["bias"] = kwargs.get("bias", norm_layer is None)
    layers: List[nn.Module] = [conv_layer(in_channels, out_channels, **kwargs)]
    if callable(norm_layer):
        layers.append(norm_layer(bn_channels))
    if callable(act_layer):
        layers.append(act_layer)
    if callable(attention_layer):
        layers.append(attention_layer(bn_channels))
    if callable(drop_layer):
        layers.append(drop_layer(inplace=True))
    return layers


def fuse_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> Tuple[torch.Tensor, torch.Tensor]:
    """Folds an (eval-mode) BatchNorm into the preceding convolution: returns the fused kernel and bias.

    k' = gamma / sqrt(running_var + eps) * k ;  b' = beta - gamma * running_mean / sqrt(running_var + eps) (+ scaled conv bias).
    Weight-sized host-side arithmetic in fp32 (one-off at re-parametrisation time), same operation order as the
    reference so the re-parametrised logits keep their argmax (BASELINE.json config 1).
    """
    if bn.bias.data.shape[0] != conv.weight.data.shape[0]:
        raise AssertionError("expected same number of output channels for both `conv` and `bn`")
    scale = bn.weight.data / torch.sqrt(bn.running_var + bn.eps)
    fused_bias = bn.bias.data - scale * bn.running_mean
    if conv.bias is not None:
        logger.warning("convolution layers placed before batch normalization should not have a bias.")
        fused_bias += scale * conv.bias.data
    fused_kernel = scale.view(-1, 1, 1, 1) * conv.weight.data
    return fused_kernel, fused_bias
