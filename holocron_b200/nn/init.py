"""Parameter initialisation — mirrors holocron/nn/init.py:10-24 (host-side, one-off; determines seed parity)."""
from torch import nn


def init_module(module: nn.Module, nonlinearity: str = "relu") -> None:
    """Kaiming-normal (fan_out) for convolutions, zero conv bias, BatchNorm weight=1 / bias=0.

    ``nn.Linear`` layers keep the PyTorch default, as in the reference. The RNG calls happen in ``module.modules()``
    order, so a model built under ``torch.manual_seed(s)`` gets the same parameters as the reference model.
    """
    for m in module.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity=nonlinearity)
            if m.bias is not None:
                m.bias.data.zero_()
        elif isinstance(m, nn.BatchNorm2d):
            m.weight.data.fill_(1.0)
            m.bias.data.zero_()
