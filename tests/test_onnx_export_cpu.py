"""ONNX export (reference scripts/export_to_onnx.py, tests/test_models_classification.py:116-139) on the CPU.
Every architecture family of the zoo is exported in inference form (RepVGG / MobileOne re-parametrised first, like the
reference's script), the file is parsed back and executed operator by operator (tests/_onnx_interp.py) on the seeded fixture
inputs, and the result must match the eval-mode logits recorded from the UNMODIFIED reference (tests/golden/zoo*.pt,
frozen-BatchNorm fixtures == eval-mode forward) to fp32 round-off."""
import pytest
import torch

import holocron_b200 as hb
from holocron_b200.onnx import export_onnx, proto as P

import _conditioning as C
import _onnx_interp as ORT
from conftest import load_golden


def build(name):
    torch.manual_seed(0)
    m = getattr(hb.models, name)(num_classes=10)
    return C.condition(m).eval()


def golden(name):
    f = "zoo_resnet" if name in C.CLS_RESNET else "zoo_f3" if name in C.CLS_F3 else "zoo"
    return load_golden(f)[name]["eval"]["logits"]


@pytest.mark.parametrize("name", ["repvgg_a0", "mobileone_s0", "rexnet1_0x", "darknet24", "darknet19", "darknet53",
                                  "cspdarknet53_mish", "resnet18", "resnet50d", "resnext50_32x4d", "res2net50_26w_4s", "sknet50",
                                  "convnext_atto"])
def test_exported_graph_reproduces_reference_logits(name, tmp_path):
    m = build(name)
    if hasattr(m, "reparametrize"):
        m.reparametrize()
    x, _ = C.cls_inputs(name, "eval")
    path = tmp_path / f"{name}.onnx"
    data = export_onnx(m, x, path, opset_version=14)
    assert path.read_bytes() == data
    model = ORT.load(data)
    assert model.opset_import[0].version == 14 and model.opset_import[0].domain == ""
    ops = {n.op_type for n in model.graph.node}
    assert "Conv" in ops and ("Gemm" in ops or name == "darknet19")          # Darknet-19's classifier is a 1x1 convolution
    if hasattr(m, "reparametrize"):
        assert "BatchNormalization" not in ops          # folded by the re-parametrisation
    out = ORT.run(model, x)
    ref = golden(name)
    err = ((out - ref).norm() / ref.norm()).item()
    assert out.shape == ref.shape and err < 5e-4, err
    # every value a node reads is produced before it (topological order), every initializer is used
    seen = {model.graph.input[0].name, ""} | {t.name for t in model.graph.initializer}
    used = set()
    for n in model.graph.node:
        assert all(i in seen for i in n.input), (n.op_type, list(n.input))
        used.update(n.input)
        seen.update(n.output)
    assert all(t.name in used for t in model.graph.initializer)
    # parameters travel as raw little-endian fp32 (what `export_params=True` writes)
    n_params = sum(p.numel() for p in m.parameters())
    n_init = sum(int(torch.tensor(list(t.dims)).prod()) if t.dims else 1 for t in model.graph.initializer if t.data_type == P.DT_FLOAT)
    assert n_init >= n_params


def test_export_refuses_training_form_and_writes_static_shapes(tmp_path):
    m = hb.models.repvgg_a0(num_classes=10)
    with pytest.raises(RuntimeError):
        export_onnx(m.train(), (1, 3, 64, 64))
    m.eval()
    data = export_onnx(m, (2, 3, 64, 64))          # train-form blocks in eval mode: three BatchNorm'd branches per block
    model = ORT.load(data)
    assert [d.dim_value for d in model.graph.input[0].type.tensor_type.shape.dim] == [2, 3, 64, 64]
    assert [d.dim_value for d in model.graph.output[0].type.tensor_type.shape.dim] == [2, 10]
    assert sum(n.op_type == "BatchNormalization" for n in model.graph.node) > 40
    with pytest.raises(ValueError):
        export_onnx(m, (1, 3, 64, 64), opset_version=9)
