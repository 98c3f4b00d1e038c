"""Export-time description of the fused entry points in ATen terms.

The module trees of :mod:`holocron_b200.models` call the CUDA entry points of :mod:`holocron_b200.nn._fused`, which a tracer
cannot look into. For export (and only there) every entry point is swapped for its definition as stock torch operators in
INFERENCE form, and the model's ``forward`` is traced over *fake* tensors (``torch.fx.experimental.proxy_tensor.make_fx``,
``tracing_mode="fake"``): shapes propagate, nothing is computed - this is a graph description, not an execution path (the
product path stays CUDA-only). Training-mode BatchNorm is refused: exported graphs are inference graphs, like the reference's
(scripts/export_to_onnx.py puts the model in ``eval()`` and re-parametrises RepVGG / MobileOne first)."""
import contextlib
import importlib
from typing import Optional, Sequence

import torch
import torch.nn.functional as TF
from torch import Tensor, nn


def _act(z: Tensor, code: int, slope: float) -> Tensor:
    if code == 1:
        return torch.relu(z)
    if code == 2:
        return torch.clamp(z, 0.0, 6.0)
    if code == 3:
        return z * torch.sigmoid(z)
    if code == 4:
        return TF.leaky_relu(z, slope)
    if code == 5:
        return z * torch.tanh(TF.softplus(z))
    if code == 6:
        return 0.5 * z * torch.clamp(z + 2, 0.0, 2.0)
    if code == 0:
        return z
    raise NotImplementedError(f"activation code {code} has no export lowering")


def _narrow(t: Tensor, c: int) -> Tensor:
    return t if t.shape[1] == c else t[:, :c]


def _bn_eval(u: Tensor, bn: nn.BatchNorm2d) -> Tensor:
    if bn.training or bn.running_mean is None:
        raise RuntimeError("export needs inference-mode BatchNorm: call model.eval() (and reparametrize()) first")
    return TF.batch_norm(_narrow(u, bn.num_features), bn.running_mean, bn.running_var, bn.weight, bn.bias, False, 0.0, bn.eps)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, keep_padded=False, want_stats=False):
    return TF.conv2d(_narrow(x, weight.shape[1]), weight, bias, stride, padding, dilation)


def conv2d_bias_act(x, weight, bias, stride, padding, act=0, slope=0.0):
    return _act(conv2d(x, weight, bias, stride, padding), act, slope)


def bn_act(us: Sequence[Tensor], bns: Sequence[nn.BatchNorm2d], act: int = 0, slope: float = 0.0,
           residual: Optional[Tensor] = None, training: Optional[bool] = None, res_after_act: bool = False,
           emit_stats: bool = False) -> Tensor:
    z = None
    for u, bn in zip(us, bns):
        t = _bn_eval(u, bn)
        z = t if z is None else z + t
    if residual is not None and not res_after_act:
        r = residual
        if r.shape[1] < z.shape[1]:                       # partial-channel shortcut (ReXNet): zero-extend
            r = TF.pad(r, (0, 0, 0, 0, 0, z.shape[1] - r.shape[1]))
        r = _narrow(r, z.shape[1])
        z = torch.maximum(z, r) if act == 7 else z + r
    z = z if act == 7 else _act(z, act, slope)
    if residual is not None and res_after_act:
        z = z + _narrow(residual, z.shape[1])
    return z


def act_only(x, act, slope=0.0):
    return _act(x, act, slope)


def gate_act(x, gate, act=0, slope=0.0):
    return _act(x * _narrow(gate, x.shape[1]), act, slope)


def repblock(x, w3, w1, bns, stride, act, slope, training):
    if training:
        raise RuntimeError("export needs inference-mode BatchNorm: call model.eval() (and reparametrize()) first")
    xf = _narrow(x, w3.shape[1])
    z = _bn_eval(TF.conv2d(xf, w3, None, stride, 1), bns[0]) + _bn_eval(TF.conv2d(xf, w1, None, stride, 0), bns[1])
    if len(bns) == 3:
        z = z + _bn_eval(xf, bns[2])
    return _act(z, act, slope)


def to_channels_last_bf16(x, c_pad=None):
    return x


def global_avg_pool_flat(x):
    return x.mean((2, 3))


def head_linear(feats, weight, bias):
    return TF.linear(feats, weight, bias)


def dwconv2d(x, weight, bias=None, stride=1, padding=0):
    c = weight.shape[0]
    return TF.conv2d(_narrow(x, c), weight, bias, stride, padding, 1, c)


@contextlib.contextmanager
def lowered():
    """Swaps the fused entry points (and the names bound to them in the model files) for the definitions above."""
    fused = importlib.import_module("holocron_b200.nn._fused")
    dw = importlib.import_module("holocron_b200.nn._dwconv")
    rexnet = importlib.import_module("holocron_b200.models.classification.rexnet")
    swaps = [(fused, n, globals()[n]) for n in ("conv2d", "conv2d_bias_act", "bn_act", "act_only", "gate_act", "repblock",
                                                "to_channels_last_bf16", "global_avg_pool_flat", "head_linear")]
    swaps += [(dw, "dwconv2d", dwconv2d), (rexnet, "dwconv2d", dwconv2d)]
    saved = [(m, n, getattr(m, n)) for m, n, _ in swaps]
    try:
        for m, n, fn in swaps:
            setattr(m, n, fn)
        yield
    finally:
        for m, n, fn in saved:
            setattr(m, n, fn)
