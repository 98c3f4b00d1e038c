"""ONNX export of inference-form models (reference scripts/export_to_onnx.py and
tests/test_models_classification.py:116-139: ``torch.onnx.export(model.eval(), rand(1, 3, H, W), path, opset_version=14)``,
RepVGG / MobileOne re-parametrised first).

``torch.onnx`` needs the ``onnx`` package (absent here) and could not look into the CUDA entry points anyway. The exporter:
  1. swaps the fused entry points for their ATen definitions (:mod:`._lowering`) and traces ``model(x)`` over FAKE tensors with
     ``make_fx`` - an ATen-level graph, parameters as graph constants, no computation;
  2. drops everything the output does not depend on, and maps each ATen node onto opset-14 operators (table below; LayerNorm
     and GELU, which opset 14 lacks, are expanded into primitives like ``torch.onnx`` does);
  3. serialises ``ModelProto`` with the schema subset of :mod:`.proto`.
Shapes are static (those of the sample input), like the reference's export. ``tests/test_onnx_export_cpu.py`` reads the files
back and executes them node by node against the reference's eval-mode logits."""
import math
import operator
from pathlib import Path
from typing import Any, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from torch import Tensor, nn

from . import proto as P
from ._lowering import lowered

__all__ = ["export_onnx"]

aten = torch.ops.aten


class _Graph:
    def __init__(self) -> None:
        self.nodes: List[Any] = []
        self.inits: Dict[str, np.ndarray] = {}
        self._n = 0

    def fresh(self, hint: str = "t") -> str:
        self._n += 1
        return f"{hint}_{self._n}"

    def const(self, value: Union[np.ndarray, Sequence, float, int], dtype=np.float32, hint: str = "c") -> str:
        name = self.fresh(hint)
        self.inits[name] = np.asarray(value, dtype=dtype)
        return name

    def node(self, op: str, inputs: Sequence[str], n_out: int = 1, outputs: Optional[Sequence[str]] = None, **attrs) -> List[str]:
        outs = list(outputs) if outputs is not None else [self.fresh(op.lower()) for _ in range(n_out)]
        n = P.NodeProto()
        n.op_type, n.name = op, self.fresh(f"n_{op}")
        n.input.extend(inputs)
        n.output.extend(outs)
        for k, v in attrs.items():
            a = n.attribute.add()
            a.name = k
            if isinstance(v, float):
                a.type, a.f = P.ATTR_FLOAT, v
            elif isinstance(v, (bool, int)):
                a.type, a.i = P.ATTR_INT, int(v)
            elif isinstance(v, str):
                a.type, a.s = P.ATTR_STRING, v.encode()
            elif isinstance(v, (list, tuple)) and all(isinstance(e, (bool, int)) for e in v):
                a.type = P.ATTR_INTS
                a.ints.extend(int(e) for e in v)
            elif isinstance(v, (list, tuple)):
                a.type = P.ATTR_FLOATS
                a.floats.extend(float(e) for e in v)
            else:
                raise TypeError(f"attribute {k}={v!r}")
        self.nodes.append(n)
        return outs


def _pair(v) -> List[int]:
    v = list(v) if isinstance(v, (list, tuple)) else [v]
    return [int(v[0])] * 2 if len(v) == 1 else [int(e) for e in v]


def _tensor_proto(name: str, arr: np.ndarray):
    t = P.TensorProto()
    t.name = name
    t.dims.extend(arr.shape)
    if arr.dtype == np.float32:
        t.data_type = P.DT_FLOAT
    elif arr.dtype == np.int64:
        t.data_type = P.DT_INT64
    else:
        raise TypeError(f"initializer dtype {arr.dtype}")
    t.raw_data = np.ascontiguousarray(arr).tobytes()      # little-endian, row-major
    return t


def _value_info(name: str, shape: Sequence[int]):
    v = P.ValueInfoProto()
    v.name = name
    v.type.tensor_type.elem_type = P.DT_FLOAT
    for d in shape:
        v.type.tensor_type.shape.dim.add().dim_value = int(d)
    return v


def _live_nodes(gm: torch.fx.GraphModule) -> List[torch.fx.Node]:
    out = next(n for n in gm.graph.nodes if n.op == "output")
    live, stack = set(), [out]
    while stack:
        n = stack.pop()
        if n in live:
            continue
        live.add(n)
        stack.extend(n.all_input_nodes)
    return [n for n in gm.graph.nodes if n in live]


def _shape(n: torch.fx.Node) -> Tuple[int, ...]:
    v = n.meta["val"]
    return tuple(int(d) for d in v.shape)


def _convert(gm: torch.fx.GraphModule, input_name: str, output_name: str):
    g = _Graph()
    env: Dict[torch.fx.Node, Any] = {}        # fx node -> value name (or list of names for multi-output nodes)

    def val(a) -> str:
        if isinstance(a, torch.fx.Node):
            v = env[a]
            if isinstance(v, list):
                raise RuntimeError(f"multi-output node {a} used without getitem")
            return v
        if isinstance(a, (int, float, bool)):
            return g.const(float(a))
        raise TypeError(f"unsupported operand {a!r}")

    def unary(op):
        return lambda n, x, *a, **k: g.node(op, [val(x)])[0]

    def binary(op):
        def f(n, a, b, alpha=1, **k):
            if alpha != 1:
                raise NotImplementedError("alpha != 1")
            return g.node(op, [val(a), val(b)])[0]
        return f

    def conv(n, x, w, b, stride, padding, dilation, transposed, output_padding, groups):
        if transposed:
            raise NotImplementedError("transposed convolution")
        p = _pair(padding)
        ins = [val(x), val(w)] + ([val(b)] if b is not None else [])
        return g.node("Conv", ins, strides=_pair(stride), pads=p + p, dilations=_pair(dilation), group=int(groups),
                      kernel_shape=list(_shape(w)[2:]))[0]

    def batch_norm(n, x, w, b, rm, rv, training, momentum, eps):
        if training:
            raise RuntimeError("training-mode BatchNorm in an export graph")
        c = _shape(x)[1]
        scale = val(w) if w is not None else g.const(np.ones(c))
        shift = val(b) if b is not None else g.const(np.zeros(c))
        return [g.node("BatchNormalization", [val(x), scale, shift, val(rm), val(rv)], epsilon=float(eps))[0], None, None]

    def layer_norm(n, x, normalized_shape, w, b, eps):
        axes = list(range(-len(normalized_shape), 0))
        xv = val(x)
        mean = g.node("ReduceMean", [xv], axes=axes, keepdims=1)[0]
        d = g.node("Sub", [xv, mean])[0]
        var = g.node("ReduceMean", [g.node("Mul", [d, d])[0]], axes=axes, keepdims=1)[0]
        std = g.node("Sqrt", [g.node("Add", [var, g.const(float(eps))])[0]])[0]
        y = g.node("Div", [d, std])[0]
        if w is not None:
            y = g.node("Mul", [y, val(w)])[0]
        if b is not None:
            y = g.node("Add", [y, val(b)])[0]
        return [y, None, None]

    def gelu(n, x, approximate="none"):
        if approximate != "none":
            raise NotImplementedError("tanh-approximated GELU")
        xv = val(x)
        e = g.node("Erf", [g.node("Mul", [xv, g.const(1.0 / math.sqrt(2.0))])[0]])[0]
        return g.node("Mul", [g.node("Mul", [xv, g.const(0.5)])[0], g.node("Add", [e, g.const(1.0)])[0]])[0]

    def clip(n, x, lo=None, hi=None):
        ins = [val(x), g.const(float(lo)) if lo is not None else "", g.const(float(hi)) if hi is not None else ""]
        while ins and ins[-1] == "":
            ins.pop()
        return g.node("Clip", ins)[0]

    def softplus(n, x, beta=1, threshold=20):
        if beta != 1:
            raise NotImplementedError("softplus beta != 1")
        return g.node("Softplus", [val(x)])[0]

    def pool(op):
        def f(n, x, kernel, stride=(), padding=0, *rest, **kw):
            k, p = _pair(kernel), _pair(padding)
            s = _pair(stride) if stride not in ((), [], None) else k
            attrs = dict(kernel_shape=k, strides=s, pads=p + p)
            if op == "MaxPool":
                dilation = rest[0] if len(rest) > 0 else 1
                ceil = rest[1] if len(rest) > 1 else False
                attrs.update(dilations=_pair(dilation), ceil_mode=int(bool(ceil)))
                return [g.node("MaxPool", [val(x)], **attrs)[0], None]
            ceil = rest[0] if len(rest) > 0 else False
            include_pad = rest[1] if len(rest) > 1 else True
            if len(rest) > 2 and rest[2] is not None:
                raise NotImplementedError("avg_pool2d divisor_override")
            attrs.update(ceil_mode=int(bool(ceil)), count_include_pad=int(bool(include_pad)))
            return g.node("AveragePool", [val(x)], **attrs)[0]
        return f

    def mean(n, x, dims, keepdim=False, **kw):
        return g.node("ReduceMean", [val(x)], axes=[int(d) for d in dims], keepdims=int(bool(keepdim)))[0]

    def reshape(n, x, shape):
        return g.node("Reshape", [val(x), g.const([int(s) for s in _shape(n)], np.int64, "shape")])[0]

    def cat(n, tensors, dim=0):
        return g.node("Concat", [val(t) for t in tensors], axis=int(dim))[0]

    def split(n, x, size, dim=0):
        total = _shape(x)[dim]
        sizes = [size] * (total // size) + ([total % size] if total % size else [])
        return g.node("Split", [val(x), g.const(sizes, np.int64, "split")], n_out=len(sizes), axis=int(dim))

    def select(n, x, dim, index):
        return g.node("Gather", [val(x), g.const(int(index), np.int64, "idx")], axis=int(dim))[0]

    def slice_(n, x, dim=0, start=None, end=None, step=1):
        size = _shape(x)[dim]
        start = 0 if start is None else int(start)
        end = size if end is None else min(int(end), size)
        if start == 0 and end == size and step == 1:
            return val(x)
        return g.node("Slice", [val(x), g.const([start], np.int64), g.const([end], np.int64), g.const([int(dim)], np.int64),
                                g.const([int(step)], np.int64)])[0]

    def pad(n, x, pads, value=0.0):
        rank = len(_shape(x))
        begins, ends = [0] * rank, [0] * rank
        for i in range(len(pads) // 2):                     # torch: last dimension first, (before, after) pairs
            begins[rank - 1 - i], ends[rank - 1 - i] = int(pads[2 * i]), int(pads[2 * i + 1])
        return g.node("Pad", [val(x), g.const(begins + ends, np.int64, "pads"), g.const(float(value))], mode="constant")[0]

    def addmm(n, bias, a, b, beta=1, alpha=1):
        return g.node("Gemm", [val(a), val(b), val(bias)], alpha=float(alpha), beta=float(beta))[0]

    def softmax(n, x, dim, half_to_float=False):
        return g.node("Softmax", [val(x)], axis=int(dim))[0]

    def upsample(mode):
        def f(n, x, output_size, *rest, **kw):
            align = bool(rest[0]) if (mode == "linear" and rest) else False
            sizes = g.const(list(_shape(n)), np.int64, "sizes")
            attrs = dict(mode=mode, coordinate_transformation_mode="align_corners" if align else
                         ("asymmetric" if mode == "nearest" else "half_pixel"))
            if mode == "nearest":
                attrs["nearest_mode"] = "floor"
            return g.node("Resize", [val(x), "", "", sizes], **attrs)[0]
        return f

    identity = lambda n, x, *a, **k: val(x)   # noqa: E731  (alias / clone / dtype-preserving copies / eval-mode dropout)

    table = {
        aten.convolution.default: conv, aten.native_batch_norm.default: batch_norm,
        aten._native_batch_norm_legit_no_training.default:
            lambda n, x, w, b, rm, rv, momentum, eps: batch_norm(n, x, w, b, rm, rv, False, momentum, eps),
        aten.native_layer_norm.default: layer_norm, aten.gelu.default: gelu,
        aten.relu.default: unary("Relu"), aten.relu_.default: unary("Relu"), aten.sigmoid.default: unary("Sigmoid"),
        aten.tanh.default: unary("Tanh"), aten.softplus.default: softplus, aten.exp.default: unary("Exp"),
        aten.sqrt.default: unary("Sqrt"), aten.erf.default: unary("Erf"),
        aten.leaky_relu.default: lambda n, x, slope=0.01: g.node("LeakyRelu", [val(x)], alpha=float(slope))[0],
        aten.leaky_relu_.default: lambda n, x, slope=0.01: g.node("LeakyRelu", [val(x)], alpha=float(slope))[0],
        aten.hardtanh.default: clip, aten.hardtanh_.default: clip, aten.clamp.default: clip,
        aten.add.Tensor: binary("Add"), aten.sub.Tensor: binary("Sub"), aten.mul.Tensor: binary("Mul"),
        aten.div.Tensor: binary("Div"), aten.maximum.default: binary("Max"), aten.add_.Tensor: binary("Add"),
        aten.max_pool2d_with_indices.default: pool("MaxPool"), aten.avg_pool2d.default: pool("AveragePool"),
        aten.mean.dim: mean, aten.view.default: reshape, aten._unsafe_view.default: reshape, aten.reshape.default: reshape,
        aten.permute.default: lambda n, x, dims: g.node("Transpose", [val(x)], perm=[int(d) for d in dims])[0],
        aten.t.default: lambda n, x: g.node("Transpose", [val(x)], perm=[1, 0])[0],
        aten.cat.default: cat, aten.split.Tensor: split, aten.select.int: select, aten.slice.Tensor: slice_,
        aten.constant_pad_nd.default: pad, aten.addmm.default: addmm, aten._softmax.default: softmax,
        aten.mm.default: lambda n, a, b: g.node("MatMul", [val(a), val(b)])[0],
        aten.upsample_nearest2d.default: upsample("nearest"), aten.upsample_bilinear2d.default: upsample("linear"),
        aten.alias.default: identity, aten.clone.default: identity, aten.detach.default: identity,
        aten._to_copy.default: identity, aten.contiguous.default: identity, aten.dropout.default: identity,
    }

    for n in _live_nodes(gm):
        if n.op == "placeholder":
            env[n] = input_name
        elif n.op == "get_attr":
            t = getattr(gm, n.target)
            if not isinstance(t, Tensor):
                raise TypeError(f"constant {n.target} is not a tensor")
            t = t.detach().cpu()
            env[n] = g.const(t.numpy().astype(np.int64) if not t.is_floating_point() else t.float().numpy(),
                             np.int64 if not t.is_floating_point() else np.float32, "w")
        elif n.op == "call_function":
            if n.target is operator.getitem:
                src, idx = n.args
                env[n] = env[src][idx]
                if env[n] is None:
                    raise NotImplementedError(f"output {idx} of {src.target} (training statistics / pooling indices)")
                continue
            fn = table.get(n.target)
            if fn is None:
                raise NotImplementedError(f"no ONNX mapping for {n.target} (used at {n.stack_trace or n.name})")
            env[n] = fn(n, *n.args, **n.kwargs)
        elif n.op == "output":
            res = n.args[0]
            res = res[0] if isinstance(res, (tuple, list)) else res
            g.node("Identity", [val(res)], outputs=[output_name])
            out_shape = _shape(res)
    return g, out_shape


def export_onnx(model: nn.Module, sample: Union[Tensor, Sequence[int]], path: Optional[Union[str, Path]] = None,
                opset_version: int = 14, input_name: str = "input", output_name: str = "output") -> bytes:
    """Serialises ``model`` (inference form: ``model.eval()``, re-parametrised where the architecture offers it) for a
    static input shape. ``sample``: a tensor or a shape, e.g. ``(1, 3, 224, 224)``. Returns the file's bytes (and writes them
    to ``path`` when given)."""
    from torch.fx.experimental.proxy_tensor import make_fx
    if opset_version < 13 or opset_version > 17:
        raise ValueError("the operator table targets opsets 13 to 17 (the reference exports opset 14)")
    if model.training:
        raise RuntimeError("export needs an inference-form model: call model.eval() (and reparametrize()) first")
    shape = tuple(sample.shape) if isinstance(sample, Tensor) else tuple(int(s) for s in sample)
    x = torch.zeros(shape, dtype=torch.float32)
    cpu_model = model if all(p.device.type == "cpu" for p in model.parameters()) else None
    if cpu_model is None:
        import copy
        cpu_model = copy.deepcopy(model).cpu()
    with lowered(), torch.no_grad():
        gm = make_fx(lambda t: cpu_model(t), tracing_mode="fake", _allow_non_fake_inputs=True)(x)
    g, out_shape = _convert(gm, input_name, output_name)
    m = P.ModelProto()
    m.ir_version = 7
    m.producer_name, m.producer_version = "holocron_b200", "0.1"
    op = m.opset_import.add()
    op.domain, op.version = "", int(opset_version)
    m.graph.name = type(model).__name__
    m.graph.node.extend(g.nodes)
    m.graph.initializer.extend(_tensor_proto(k, v) for k, v in g.inits.items())
    m.graph.input.append(_value_info(input_name, shape))
    m.graph.output.append(_value_info(output_name, out_shape))
    data = m.SerializeToString()
    if path is not None:
        Path(path).write_bytes(data)
    return data
