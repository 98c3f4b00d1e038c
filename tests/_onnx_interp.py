"""A small ONNX executor for the tests (the ``onnx`` / ``onnxruntime`` packages are not in this image): parses a serialised
``ModelProto`` with the schema subset of ``holocron_b200.onnx.proto`` and evaluates the graph node by node with torch fp32 ops
following the operator specifications (opset 13-17 semantics of the operators the exporter emits). Test infrastructure."""
import numpy as np
import torch
import torch.nn.functional as TF

from holocron_b200.onnx import proto as P


def _attrs(node):
    out = {}
    for a in node.attribute:
        if a.type == P.ATTR_FLOAT:
            out[a.name] = a.f
        elif a.type == P.ATTR_INT:
            out[a.name] = a.i
        elif a.type == P.ATTR_STRING:
            out[a.name] = a.s.decode()
        elif a.type == P.ATTR_INTS:
            out[a.name] = list(a.ints)
        elif a.type == P.ATTR_FLOATS:
            out[a.name] = list(a.floats)
        else:
            raise NotImplementedError(f"attribute type {a.type}")
    return out


def _init(t):
    dt = {P.DT_FLOAT: np.float32, P.DT_INT64: np.int64}[t.data_type]
    return torch.from_numpy(np.frombuffer(t.raw_data, dtype=dt).reshape(tuple(t.dims)).copy())


def load(data: bytes):
    m = P.ModelProto()
    m.ParseFromString(data)
    return m


def _sym_pads(pads):
    half = len(pads) // 2
    assert list(pads[:half]) == list(pads[half:]), "asymmetric pads"
    return list(pads[:half])


def run(model, x: torch.Tensor) -> torch.Tensor:
    g = model.graph
    assert model.ir_version == 7 and len(g.input) == 1 and len(g.output) == 1
    assert tuple(d.dim_value for d in g.input[0].type.tensor_type.shape.dim) == tuple(x.shape)
    env = {t.name: _init(t) for t in g.initializer}
    env[g.input[0].name] = x.float()
    env[""] = None
    for node in g.node:
        a = _attrs(node)
        i = [env[n] for n in node.input]
        op = node.op_type
        if op == "Conv":
            y = TF.conv2d(i[0], i[1], i[2] if len(i) > 2 else None, a["strides"], _sym_pads(a["pads"]), a["dilations"], a["group"])
            assert list(i[1].shape[2:]) == a["kernel_shape"]
        elif op == "BatchNormalization":
            y = TF.batch_norm(i[0], i[3], i[4], i[1], i[2], False, 0.0, a["epsilon"])
        elif op == "Relu":
            y = torch.relu(i[0])
        elif op == "LeakyRelu":
            y = TF.leaky_relu(i[0], a["alpha"])
        elif op == "Clip":
            lo = i[1].item() if len(i) > 1 and i[1] is not None else None
            hi = i[2].item() if len(i) > 2 and i[2] is not None else None
            y = torch.clamp(i[0], lo, hi)
        elif op in ("Sigmoid", "Tanh", "Erf", "Sqrt", "Exp"):
            y = getattr(torch, op.lower())(i[0])
        elif op == "Softplus":
            y = TF.softplus(i[0])
        elif op in ("Add", "Sub", "Mul", "Div"):
            y = {"Add": torch.add, "Sub": torch.sub, "Mul": torch.mul, "Div": torch.div}[op](i[0], i[1])
        elif op == "Max":
            y = torch.maximum(i[0], i[1])
        elif op == "MaxPool":
            y = TF.max_pool2d(i[0], a["kernel_shape"], a["strides"], _sym_pads(a["pads"]), a.get("dilations", 1), bool(a.get("ceil_mode", 0)))
        elif op == "AveragePool":
            y = TF.avg_pool2d(i[0], a["kernel_shape"], a["strides"], _sym_pads(a["pads"]), bool(a.get("ceil_mode", 0)),
                              bool(a.get("count_include_pad", 0)))
        elif op == "ReduceMean":
            y = i[0].mean(a["axes"], keepdim=bool(a.get("keepdims", 1)))
        elif op == "Reshape":
            y = i[0].reshape([int(s) for s in i[1]])
        elif op == "Transpose":
            y = i[0].permute(a["perm"])
        elif op == "Concat":
            y = torch.cat(i, a["axis"])
        elif op == "Split":
            y = list(torch.split(i[0], [int(s) for s in i[1]], a.get("axis", 0)))
        elif op == "Gather":
            assert i[1].ndim == 0
            y = i[0].select(a.get("axis", 0), int(i[1]))
        elif op == "Slice":
            y = i[0]
            for s, e, ax, st in zip(i[1].tolist(), i[2].tolist(), i[3].tolist(), i[4].tolist()):
                y = y.narrow(ax, s, e - s)[(slice(None),) * ax + (slice(None, None, st),)]
        elif op == "Pad":
            pads = i[1].tolist()
            rank = len(pads) // 2
            flat = []
            for d in reversed(range(rank)):
                flat += [pads[d], pads[rank + d]]
            y = TF.pad(i[0], flat, value=float(i[2]) if len(i) > 2 and i[2] is not None else 0.0)
            assert a.get("mode", "constant") == "constant"
        elif op == "Gemm":
            assert not a.get("transA", 0) and not a.get("transB", 0)
            y = a.get("alpha", 1.0) * (i[0] @ i[1]) + a.get("beta", 1.0) * i[2]
        elif op == "MatMul":
            y = i[0] @ i[1]
        elif op == "Softmax":
            y = torch.softmax(i[0], a.get("axis", -1))
        elif op == "Resize":
            size = [int(s) for s in i[3]][2:]
            if a["mode"] == "nearest":
                assert a["coordinate_transformation_mode"] == "asymmetric" and a["nearest_mode"] == "floor"
                y = TF.interpolate(i[0], size=size, mode="nearest")
            else:
                y = TF.interpolate(i[0], size=size, mode="bilinear",
                                   align_corners=a["coordinate_transformation_mode"] == "align_corners")
        elif op == "Identity":
            y = i[0]
        else:
            raise NotImplementedError(op)
        ys = y if isinstance(y, list) else [y]
        assert len(ys) == len(node.output)
        for name, v in zip(node.output, ys):
            env[name] = v
    out = env[g.output[0].name]
    assert tuple(d.dim_value for d in g.output[0].type.tensor_type.shape.dim) == tuple(out.shape)
    return out
