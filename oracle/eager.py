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

``reference_execution(bf16_storage=True)`` additionally rounds every tensor to bfloat16 exactly where the CUDA path
STORES one (convolution outputs, fused BatchNorm/activation outputs, pooled features, classifier GEMM operands and
result; filters are rounded where the kernels read their bf16 packed copy) while all arithmetic stays fp32, like the
kernels' fp32 accumulators. A deep random-init network in training mode amplifies independent bf16 rounding noise by
~1.2x per BatchNorm layer (the batch mean it removes carries signal energy but no noise), so ANY bf16 execution lands
0.05 - 0.5 rel-L2 from the fp32 fixture; with the SAME storage points the two executions make the same rounding decisions
(up to fp32 summation order) and stay within ~1e-2 of each other through 50+ layers - which is what
``tests/test_gpu_zoo.py`` asserts end to end for the training-mode forward of every zoo model. The fp32 mode of this very
executor is pinned to the reference's fixtures (2e-4), so the chain CUDA == bf16-storage executor, fp32 executor ==
reference is tight on both links.
"""
import contextlib
from typing import Optional, Sequence

import torch
import torch.nn.functional as TF
from torch import Tensor, nn


_BF16 = False   # set by reference_execution(bf16_storage=True)


class _RoundBF16(torch.autograd.Function):
    """Round-to-nearest-even to bfloat16 (values stay fp32 tensors); the gradient is rounded the same way, like the bf16
    gradient tensors the CUDA path passes between its backward kernels."""

    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g: Tensor) -> Tensor:
        return g.to(torch.bfloat16).float()


def _st(x: Tensor) -> Tensor:
    """Storage rounding point of the CUDA path (identity in fp32 mode)."""
    return _RoundBF16.apply(x) if _BF16 else x


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


def _f(t: Tensor) -> Tensor:
    """fp32 view of an activation - except under torch autocast, where the tensor keeps the dtype autocast gave it (the
    executor then behaves like stock torch modules under bf16 autocast: tests use that 'autocast twin' as the yardstick of
    what a library bf16 execution of the same network achieves)."""
    if t.is_cuda and torch.is_autocast_enabled():
        return t
    return t.float()


def _conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0, dilation: int = 1,
            keep_padded: bool = False, want_stats: bool = False) -> Tensor:
    # bf16 mode: activations arrive rounded, the kernels read the bf16-packed filter, the bias is added in fp32
    return _st(TF.conv2d(_st(_f(_narrow(x, weight.shape[1]))), _st(weight), bias, stride, padding, dilation))


def _conv2d_bias_act(x, weight, bias, stride, padding, act=0, slope=0.0):
    y = _conv2d(x, weight, bias, stride, padding)
    return y if act == 0 else _st(_act(y, act, slope))


def _bn_act(us: Sequence[Tensor], bns: Sequence[nn.BatchNorm2d], act: int = 0, slope: float = 0.0,
            residual: Optional[Tensor] = None, training: Optional[bool] = None, res_after_act: bool = False,
            emit_stats: bool = False) -> Tensor:
    z = None
    for u, bn in zip(us, bns):
        t = bn(_f(_narrow(u, bn.num_features)))
        z = t if z is None else z + t
    if residual is not None and not res_after_act:
        r = _f(residual)
        if r.shape[1] < z.shape[1]:                       # partial-channel shortcut (ReXNet): zero-extend
            r = TF.pad(r, (0, 0, 0, 0, 0, z.shape[1] - r.shape[1]))
        r = _narrow(r, z.shape[1])
        z = torch.maximum(z, r) if act == 7 else z + r
    z = _act(z, act, slope)
    if residual is not None and res_after_act:
        z = z + _narrow(_f(residual), z.shape[1])
    return _st(z)


def _act_only(x: Tensor, act: int, slope: float = 0.0) -> Tensor:
    return _st(_act(_f(x), act, slope))


def _gate_act(x: Tensor, gate: Tensor, act: int = 0, slope: float = 0.0) -> Tensor:
    # the squeeze path hands the gate over in the activation dtype (rexnet.py SEBlock.gate: `.to(x.dtype)`)
    return _st(_act(_f(x) * _narrow(_st(_f(gate)), x.shape[1]), act, slope))


def _repblock(x: Tensor, w3: Tensor, w1: Tensor, bns: Sequence[nn.BatchNorm2d], stride: int, act: int, slope: float,
              training: bool) -> Tensor:
    xf = _st(_f(_narrow(x, w3.shape[1])))
    z = bns[0](_st(TF.conv2d(xf, _st(w3), None, stride, 1))) + bns[1](_st(TF.conv2d(xf, _st(w1), None, stride, 0)))
    if len(bns) == 3:
        z = z + bns[2](xf)
    return _st(_act(z, act, slope))


def _to_channels_last(x: Tensor, c_pad: Optional[int] = None) -> Tensor:
    return _st(_f(x))


def _gap(x: Tensor) -> Tensor:
    return _st(_f(x).mean((2, 3)))


def _head_linear(feats: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
    return _st(TF.linear(_st(_f(feats)), _st(weight), None if bias is None else _st(bias))).float()


def _dwconv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0) -> Tensor:
    c = weight.shape[0]
    # the depth-wise kernels read the fp32 master filter directly (no bf16 packing)
    return _st(TF.conv2d(_f(_narrow(x, c)), weight, bias, stride, padding, 1, c))


@contextlib.contextmanager
def reference_execution(bf16_storage: bool = False):
    """Temporarily route the fused entry points to the torch restatements above (fp32, or fp32 arithmetic with the CUDA
    path's bf16 storage points when ``bf16_storage``)."""
    import importlib
    global _BF16

    import holocron_b200.nn._dwconv as dw
    import holocron_b200.nn._fused as fused

    from . import boxes as oracle_boxes

    rexnet = importlib.import_module("holocron_b200.models.classification.rexnet")
    yolov4 = importlib.import_module("holocron_b200.models.detection.yolov4")   # the package re-exports a function of that name
    yolo = importlib.import_module("holocron_b200.models.detection.yolo")       # _YOLO losses shared by YOLOv1 / YOLOv2

    swaps = [
        (fused, "conv2d", _conv2d), (fused, "conv2d_bias_act", _conv2d_bias_act), (fused, "bn_act", _bn_act),
        (fused, "act_only", _act_only), (fused, "gate_act", _gate_act), (fused, "repblock", _repblock),
        (fused, "to_channels_last_bf16", _to_channels_last), (fused, "global_avg_pool_flat", _gap),
        (fused, "head_linear", _head_linear),
        (dw, "dwconv2d", _dwconv2d), (rexnet, "dwconv2d", _dwconv2d),
        (yolov4, "box_iou", oracle_boxes.box_iou), (yolov4, "ciou_loss", oracle_boxes.ciou_loss),
        (yolo, "box_iou", oracle_boxes.box_iou),
    ]
    saved = [(mod, name, getattr(mod, name)) for mod, name, _ in swaps]
    prev = _BF16
    hooks = []
    try:
        _BF16 = bool(bf16_storage)
        for mod, name, fn in swaps:
            setattr(mod, name, fn)
        yield hooks
    finally:
        _BF16 = prev
        for mod, name, fn in saved:
            setattr(mod, name, fn)
        for h in hooks:
            h.remove()


def round_interpolations(model: nn.Module, hooks: list) -> None:
    """bf16-storage mode: modules the CUDA path runs as stock torch ops on bf16 tensors and whose result is NOT a pure
    selection of input values (bilinear ``nn.Upsample`` in UNet3+) store a rounded result as well. Max-pooling, nearest
    up-sampling and concatenation only move values and need no hook."""
    for m in model.modules():
        if isinstance(m, nn.Upsample) and m.mode != "nearest":
            hooks.append(m.register_forward_hook(lambda _m, _i, out: _st(out)))
