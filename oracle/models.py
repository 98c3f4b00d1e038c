"""Oracle restatements of the model-side pieces of the hot path, built from stock torch.nn layers on CPU
(reference: holocron/models/utils.py, holocron/models/classification/repvgg.py, holocron/nn/init.py).

Construction order (and therefore the RNG stream under a fixed seed) follows the reference so that
``torch.manual_seed(s); build()`` yields the same parameters as the reference model.
"""
from collections import OrderedDict
from typing import List, Tuple

import torch
from torch import Tensor, nn


def fold_conv_bn(conv_w: Tensor, gamma: Tensor, beta: Tensor, mean: Tensor, var: Tensor, eps: float) -> Tuple[Tensor, Tensor]:
    """k' = gamma / sqrt(var + eps) * k ; b' = beta - gamma * mean / sqrt(var + eps) — reference models/utils.py:116-143."""
    scale = gamma / torch.sqrt(var + eps)
    return conv_w * scale.reshape(-1, 1, 1, 1), beta - scale * mean


def init_like_reference(model: nn.Module, nonlinearity: str = "relu") -> None:
    """kaiming-normal(fan_out) for convs, zero bias, BN weight 1 / bias 0; Linear untouched — reference nn/init.py:10-24."""
    for mod in model.modules():
        if isinstance(mod, nn.Conv2d):
            nn.init.kaiming_normal_(mod.weight, mode="fan_out", nonlinearity=nonlinearity)
            if mod.bias is not None:
                mod.bias.data.zero_()
        elif isinstance(mod, nn.BatchNorm2d):
            mod.weight.data.fill_(1.0)
            mod.bias.data.zero_()


class RepBlockOracle(nn.Module):
    """relu(bn3(conv3x3(x)) + bn1(conv1x1(x)) [+ bn_id(x)]) — reference repvgg.py:38-73."""

    def __init__(self, cin: int, cout: int, stride: int, identity: bool) -> None:
        super().__init__()
        branches = [
            nn.Sequential(nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False), nn.BatchNorm2d(cout)),
            nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, padding=0, bias=False), nn.BatchNorm2d(cout)),
        ]
        if identity:
            branches.append(nn.BatchNorm2d(cout))
        self.branches = nn.ModuleList(branches)
        self.fused = None

    def forward(self, x: Tensor) -> Tensor:
        if self.fused is not None:
            return torch.relu(self.fused(x))
        out = 0
        for b in self.branches:  # python-sum order: ((0 + b3) + b1) + bid
            out = out + b(x)
        return torch.relu(out)

    @torch.no_grad()
    def reparametrize(self) -> None:
        """Fold the three branches into one 3x3 conv with bias — reference repvgg.py:75-107."""
        c3, b3 = self.branches[0]
        c1, b1 = self.branches[1]
        k3, bias3 = fold_conv_bn(c3.weight, b3.weight, b3.bias, b3.running_mean, b3.running_var, b3.eps)
        k1, bias1 = fold_conv_bn(c1.weight, b1.weight, b1.bias, b1.running_mean, b1.running_var, b1.eps)
        k = k3.clone()
        k[:, :, 1:2, 1:2] += k1
        bias = bias3 + bias1
        if len(self.branches) == 3:
            bid = self.branches[2]
            scale = bid.weight / torch.sqrt(bid.running_var + bid.eps)
            idx = torch.arange(k.shape[0])
            k[idx, idx, 1, 1] += scale
            bias = bias + bid.bias - scale * bid.running_mean
        fused = nn.Conv2d(c3.in_channels, c3.out_channels, 3, stride=c3.stride, padding=1, bias=True)
        fused.weight.data = k
        fused.bias.data = bias
        self.fused = fused


REPVGG_CFG = {
    # name: (num_blocks, planes, width multiplier a, final width multiplier b) — reference repvgg.py:206-498
    "repvgg_a0": ([1, 2, 4, 14, 1], [64, 64, 128, 256, 512], 0.75, 2.5),
    "repvgg_a1": ([1, 2, 4, 14, 1], [64, 64, 128, 256, 512], 1, 2.5),
    "repvgg_a2": ([1, 2, 4, 14, 1], [64, 64, 128, 256, 512], 1.5, 2.75),
    "repvgg_b0": ([1, 4, 6, 16, 1], [64, 64, 128, 256, 512], 1, 2.5),
}


def repvgg_channels(planes: List[int], a: float, b: float, in_channels: int = 3) -> List[int]:
    """reference repvgg.py:146-148."""
    chans = [in_channels, int(min(1, a) * planes[0])]
    chans += [int(a * c) for c in planes[1:-1]]
    chans.append(int(b * planes[-1]))
    return chans


class RepVGGOracle(nn.Sequential):
    """reference repvgg.py:110-171: 5 stages of (stride-2 block + nb identity blocks), GAP, Linear."""

    def __init__(self, name: str = "repvgg_a0", num_classes: int = 10, in_channels: int = 3) -> None:
        nb, planes, a, b = REPVGG_CFG[name]
        chans = repvgg_channels(planes, a, b, in_channels)
        stages = []
        for n, cin, cout in zip(nb, chans[:-1], chans[1:]):
            blocks = [RepBlockOracle(cin, cout, 2, False)]
            blocks += [RepBlockOracle(cout, cout, 1, True) for _ in range(n)]
            stages.append(nn.Sequential(*blocks))
        super().__init__(OrderedDict([
            ("features", nn.Sequential(*stages)),
            ("pool", nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Flatten(1))),
            ("head", nn.Linear(chans[-1], num_classes)),
        ]))
        init_like_reference(self, "relu")

    def reparametrize(self) -> None:
        for stage in self.features:
            for block in stage:
                block.reparametrize()
