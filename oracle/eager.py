"""Reference execution of the package's module trees on the CPU (TEST INFRASTRUCTURE ONLY).

The model builders of ``holocron_b200.models`` keep the reference's module trees (same children, parameter names and
init RNG order) but their ``forward`` methods call the fused CUDA entry points of ``holocron_b200.nn._fused``, which
refuse CPU tensors. Inside :func:`reference_execution` those entry points are swapped for plain fp32 torch ops with the
reference's semantics, written the way the reference writes them:

    conv -> BatchNorm2d (module call: batch statistics + running-stat update) -> activation, shortcut added where the
    reference adds it (holocron/models/utils.py:28-86, classification/resnet.py:75-87, repvgg.py:71-73, rexnet.py:131-143)

so that ``model(x)`` on the CPU runs the reference algorithm on THIS package's wiring. ``tests/test_zoo_wiring_cpu.py``
compares the result (outputs, loss, gradients) with the fixtures produced by the unmodified reference: it pins the
module trees, the channel-padding bookkeeping and the autograd wiring of every zoo model without a GPU. Nothing in the
product imports this module.
"""
import contextlib
from typing import Optional, Sequence

import torch
import torch.nn.functional as TF
from torch import Tensor, nn


def _act(z: Tensor, code: int, slope: float) -> Tensor:
    if code == 1:
        return torch.relu(z)
    if code == 2:
        return TF.relu6(z)
    if code == 3:
        return TF.silu(z)
    if code == 4:
        return TF.leaky_relu(z, slope)
    if code == 5:
        return TF.mish(z)
    if code == 6:
        return 0.5 * z * torch.clamp(z + 2, 0, 2)
    return z


def _narrow(t: Tensor, c: int) -> Tensor:
    return t if t.shape[1] == c else t[:, :c]


def _conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0, dilation: int = 1,
            keep_padded: bool = False, want_stats: bool = False) -> Tensor:
    return TF.conv2d(_narrow(x, weight.shape[1]).float(), weight, bias, stride, padding, dilation)


def _conv2d_bias_act(x, weight, bias, stride, padding, act=0, slope=0.0):
    return _act(_conv2d(x, weight, bias, stride, padding), act, slope)


def _bn_act(us: Sequence[Tensor], bns: Sequence[nn.BatchNorm2d], act: int = 0, slope: float = 0.0,
            residual: Optional[Tensor] = None, training: Optional[bool] = None, res_after_act: bool = False,
            emit_stats: bool = False) -> Tensor:
    z = None
    for u, bn in zip(us, bns):
        t = bn(_narrow(u, bn.num_features).float())
        z = t if z is None else z + t
    if residual is not None and not res_after_act:
        r = residual.float()
        if r.shape[1] < z.shape[1]:                       # partial-channel shortcut (ReXNet): zero-extend
            r = TF.pad(r, (0, 0, 0, 0, 0, z.shape[1] - r.shape[1]))
        r = _narrow(r, z.shape[1])
        z = torch.maximum(z, r) if act == 7 else z + r
    z = _act(z, act, slope)
    if residual is not None and res_after_act:
        z = z + _narrow(residual.float(), z.shape[1])
    return z


def _act_only(x: Tensor, act: int, slope: float = 0.0) -> Tensor:
    return _act(x.float(), act, slope)


def _gate_act(x: Tensor, gate: Tensor, act: int = 0, slope: float = 0.0) -> Tensor:
    return _act(x.float() * _narrow(gate.float(), x.shape[1]), act, slope)


def _repblock(x: Tensor, w3: Tensor, w1: Tensor, bns: Sequence[nn.BatchNorm2d], stride: int, act: int, slope: float,
              training: bool) -> Tensor:
    xf = _narrow(x, w3.shape[1]).float()
    z = bns[0](TF.conv2d(xf, w3, None, stride, 1)) + bns[1](TF.conv2d(xf, w1, None, stride, 0))
    if len(bns) == 3:
        z = z + bns[2](xf)
    return _act(z, act, slope)


def _to_channels_last(x: Tensor, c_pad: Optional[int] = None) -> Tensor:
    return x.float()


def _gap(x: Tensor) -> Tensor:
    return x.float().mean((2, 3))


def _dwconv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0) -> Tensor:
    c = weight.shape[0]
    return TF.conv2d(_narrow(x, c).float(), weight, bias, stride, padding, 1, c)


@contextlib.contextmanager
def reference_execution():
    """Temporarily route the fused entry points to the fp32 torch restatements above."""
    import importlib

    import holocron_b200.nn._dwconv as dw
    import holocron_b200.nn._fused as fused

    from . import boxes as oracle_boxes

    rexnet = importlib.import_module("holocron_b200.models.classification.rexnet")
    yolov4 = importlib.import_module("holocron_b200.models.detection.yolov4")   # the package re-exports a function of that name

    swaps = [
        (fused, "conv2d", _conv2d), (fused, "conv2d_bias_act", _conv2d_bias_act), (fused, "bn_act", _bn_act),
        (fused, "act_only", _act_only), (fused, "gate_act", _gate_act), (fused, "repblock", _repblock),
        (fused, "to_channels_last_bf16", _to_channels_last), (fused, "global_avg_pool_flat", _gap),
        (dw, "dwconv2d", _dwconv2d), (rexnet, "dwconv2d", _dwconv2d),
        (yolov4, "box_iou", oracle_boxes.box_iou), (yolov4, "ciou_loss", oracle_boxes.ciou_loss),
    ]
    saved = [(mod, name, getattr(mod, name)) for mod, name, _ in swaps]
    try:
        for mod, name, fn in swaps:
            setattr(mod, name, fn)
        yield
    finally:
        for mod, name, fn in saved:
            setattr(mod, name, fn)
