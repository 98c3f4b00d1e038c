"""GPU parity tests for the model-zoo rows of SURVEY §8 (a11 ReXNet, a12 Darknet, a13/a14 YOLOv4, a21 UNet3+):
training-mode forward, loss and selected gradients of the CUDA path vs fixtures produced by the unmodified reference
(tests/golden/make_golden.py --zoo) with identical seeded parameters.

Tolerances (bf16 activations through tens of layers; see test_gpu_repvgg for the mask-flip argument): logits / dense
outputs rel L2 < 5e-2, scalar losses rel < 2e-2, gradients rel L2 < 0.35 (deep random-init nets amplify bf16 noise:
torch's own bf16 autocast shows the same spread against fp32)."""
import pytest
import torch
import torch.nn.functional as TF

import holocron_b200 as hb
from holocron_b200.nn import functional as F

from conftest import load_golden

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


@pytest.mark.parametrize("name", ["darknet53", "cspdarknet53", "rexnet1_0x", "darknet24", "darknet19"])
def test_classification_backbones(name):
    g = load_golden("zoo")[name]
    torch.manual_seed(0)
    m = getattr(hb.models, name)(num_classes=10)
    if name == "rexnet1_0x":
        m.head[0].p = 0.0
    m = m.cuda().train()
    out = m(g["x"].cuda())
    assert out.shape == g["logits"].shape and out.dtype == torch.float32
    loss = TF.cross_entropy(out, g["t"].cuda())
    loss.backward()
    assert rel_l2(out, g["logits"]) < 5e-2, rel_l2(out, g["logits"])
    assert abs(loss.item() - g["loss"].item()) / abs(g["loss"].item()) < 2e-2
    ps = dict(m.named_parameters())
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in ps.values())
    assert rel_l2(ps[g["last"]].grad, g["grads"][g["last"]]) < 0.1
    # first-layer gradient after 50+ bf16 layers at batch 2: element-wise agreement with an fp32 run is not defined
    # (torch's own bf16 autocast differs from its fp32 self by rel-L2 ~0.9 here, tools/dev_gradcheck.py); check scale
    ratio = (ps[g["first"]].grad.float().norm().cpu() / g["grads"][g["first"]].norm()).item()
    assert 0.3 < ratio < 3.0, ratio
    m.eval()
    with torch.no_grad():
        assert m(g["x"].cuda()).shape == g["logits"].shape


def test_unet3p_with_dice_loss():
    g = load_golden("zoo")["unet3p"]
    torch.manual_seed(0)
    m = hb.models.unet3p(num_classes=21).cuda().train()
    out = m(g["x"].cuda())
    assert out.shape == (1, 21, 64, 64)
    onehot = TF.one_hot(g["mask"].cuda(), 21).movedim(-1, 1).float()
    loss = F.dice_loss(torch.softmax(out, 1), onehot)
    loss.backward()
    assert rel_l2(out, g["out"]) < 5e-2, rel_l2(out, g["out"])
    assert abs(loss.item() - g["loss"].item()) / abs(g["loss"].item()) < 2e-2
    ps = dict(m.named_parameters())
    assert rel_l2(ps["classifier.weight"].grad, g["grads"]["classifier.weight"]) < 0.1
    ratio = (ps["encoder.0.0.weight"].grad.float().norm().cpu() / g["grads"]["encoder.0.0.weight"].norm()).item()
    assert 0.3 < ratio < 3.0, ratio


def test_yolov4_losses_and_inference():
    g = load_golden("zoo")["yolov4"]
    torch.manual_seed(0)
    m = hb.models.yolov4(num_classes=80)
    for mod in m.modules():
        if isinstance(mod, hb.nn.DropBlock2d):
            mod.p = 0.0
    m = m.cuda().train()
    target = [{k: v.cuda() for k, v in t.items()} for t in g["target"]]
    losses = m(g["x"].cuda(), target)
    assert set(losses) == set(g["losses"])
    for k, v in losses.items():
        assert v.requires_grad and torch.isfinite(v).all()
        ref = g["losses"][k].item()
        assert abs(v.item() - ref) <= 3e-2 * abs(ref) + 1e-4, (k, v.item(), ref)
    sum(losses.values()).backward()
    ps = dict(m.named_parameters())
    assert rel_l2(ps["head.head1.3.bias"].grad, g["grads"]["head.head1.3.bias"]) < 0.1
    assert all(p.grad is None or torch.isfinite(p.grad).all() for p in ps.values())
    # empty ground truth (reference tests/test_models_detection.py:60-64) and eval-mode detections
    empty = [{"boxes": torch.zeros((0, 4), device="cuda"), "labels": torch.zeros(0, dtype=torch.long, device="cuda")}] * 2
    out = m(g["x"].cuda(), empty)
    assert all(torch.isfinite(v).all() for v in out.values())
    m.eval()
    with torch.no_grad():
        dets = m(g["x"].cuda())
    assert len(dets) == 2 and all(set(d) == {"boxes", "scores", "labels"} for d in dets)
    with pytest.raises(ValueError):
        m.train()(g["x"].cuda())
