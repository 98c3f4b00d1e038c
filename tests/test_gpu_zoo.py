"""GPU parity tests for the model-zoo rows of SURVEY §8 (a11 ReXNet, a12 Darknet, a13/a14 YOLOv4, a21 UNet3+):
training-mode forward, loss and selected gradients of the CUDA path vs fixtures produced by the unmodified reference
(tests/golden/make_golden.py --zoo) with identical seeded parameters.

Two layers of checks:
  1. teacher forcing (tests/_teacher.py): every fused launch of the forward pass is compared with fp32 torch library
     ops on the very same input tensor - conv launches, BN/activation passes and conv-BN-act units (from the
     fp32 master weights) all rel-L2 < 5e-3 (measured: 1.7e-3 = the bf16 output rounding). This is the kernel-correctness bar and it is the same for every model.
  2. end to end against the fp32 fixture with a PER-MODEL tolerance. Deep random-init nets in training mode amplify
     bf16 rounding (batch statistics over 8 samples in the last stages); the tolerance of each model is ~1.5x the
     distance at which torch's own bf16 autocast lands from the same fixture (profiles/r01_bf16_conditioning.log:
     darknet53 0.046, cspdarknet53 0.47, darknet19 0.155, darknet24 0.003), i.e. "as close as any bf16 execution".
"""
import pytest
import torch
import torch.nn.functional as TF

import holocron_b200 as hb
from holocron_b200.nn import functional as F

from _teacher import teacher_forcing
from conftest import load_golden

pytestmark = pytest.mark.gpu

# model -> (logits rel-L2, loss rel, last-layer gradient rel-L2) against the fp32 fixture
E2E_TOL = {
    "darknet53": (8e-2, 2e-2, 0.15),
    "cspdarknet53": (0.6, 5e-2, 0.6),
    "rexnet1_0x": (0.35, 2e-2, 0.35),
    "darknet24": (2e-2, 1e-2, 5e-2),
    "darknet19": (0.25, 0.25, 0.5),
}


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
    with teacher_forcing() as rep:
        out = m(g["x"].cuda())
    assert out.shape == g["logits"].shape and out.dtype == torch.float32
    loss = TF.cross_entropy(out, g["t"].cuda())
    loss.backward()
    ps = dict(m.named_parameters())
    tol_logits, tol_loss, tol_grad = E2E_TOL[name]
    e_logits = rel_l2(out, g["logits"])
    e_loss = abs(loss.item() - g["loss"].item()) / abs(g["loss"].item())
    e_grad = rel_l2(ps[g["last"]].grad, g["grads"][g["last"]])
    ratio = (ps[g["first"]].grad.float().norm().cpu() / g["grads"][g["first"]].norm()).item()
    print(f"\n[zoo] {name}: launches {rep.worst()} logits {e_logits:.4f} loss {e_loss:.4f} last-grad {e_grad:.4f} "
          f"first-grad norm ratio {ratio:.3f}")
    assert len(rep.convs) > 10
    rep.assert_ok()
    assert e_logits < tol_logits, e_logits
    assert e_loss < tol_loss, e_loss
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in ps.values())
    assert e_grad < tol_grad, e_grad
    # first-layer gradient after 50+ bf16 layers at batch 2: element-wise agreement with an fp32 run is not defined
    # (torch's own bf16 autocast differs from its fp32 self by rel-L2 ~0.9 here, tools/dev_gradcheck.py); check scale
    assert 0.3 < ratio < 3.0, ratio
    m.eval()
    with torch.no_grad():
        assert m(g["x"].cuda()).shape == g["logits"].shape


def test_unet3p_with_dice_loss():
    g = load_golden("zoo")["unet3p"]
    torch.manual_seed(0)
    m = hb.models.unet3p(num_classes=21).cuda().train()
    with teacher_forcing() as rep:
        out = m(g["x"].cuda())
    assert out.shape == (1, 21, 64, 64)
    onehot = TF.one_hot(g["mask"].cuda(), 21).movedim(-1, 1).float()
    loss = F.dice_loss(torch.softmax(out, 1), onehot)
    loss.backward()
    ps = dict(m.named_parameters())
    e_out = rel_l2(out, g["out"])
    e_loss = abs(loss.item() - g["loss"].item()) / abs(g["loss"].item())
    e_grad = rel_l2(ps["classifier.weight"].grad, g["grads"]["classifier.weight"])
    ratio = (ps["encoder.0.0.weight"].grad.float().norm().cpu() / g["grads"]["encoder.0.0.weight"].norm()).item()
    print(f"\n[zoo] unet3p: launches {rep.worst()} out {e_out:.4f} loss {e_loss:.4f} last-grad {e_grad:.4f} ratio {ratio:.3f}")
    rep.assert_ok()
    # batch 1 at 64x64: the deepest encoder stage normalises over 16 samples; bf16 autocast of the same net lands at ~0.05
    assert e_out < 0.1, e_out
    assert e_loss < 2e-2, e_loss
    assert e_grad < 0.15, e_grad
    assert 0.3 < ratio < 3.0, ratio


YOLO_TOL = {"obj_loss": 0.8, "noobj_loss": 0.8, "bbox_loss": 0.8, "clf_loss": 0.8}


def test_yolov4_losses_and_inference():
    g = load_golden("zoo")["yolov4"]
    torch.manual_seed(0)
    m = hb.models.yolov4(num_classes=80)
    for mod in m.modules():
        if isinstance(mod, hb.nn.DropBlock2d):
            mod.p = 0.0
    m = m.cuda().train()
    target = [{k: v.cuda() for k, v in t.items()} for t in g["target"]]
    with teacher_forcing() as rep:
        losses = m(g["x"].cuda(), target)
    assert set(losses) == set(g["losses"])
    print("\n[zoo] yolov4: launches", rep.worst(), {k: (round(v.item(), 4), round(g["losses"][k].item(), 4))
                                                    for k, v in losses.items()})
    rep.assert_ok()
    # The objectness / box terms depend on which anchors clear the IoU thresholds against the (random-init) predictions:
    # a discrete assignment that bf16 noise in a 100+-layer net flips for a few anchors. The per-launch checks above are
    # the parity bar; end to end the losses must stay in the fixture's neighbourhood.
    for k, v in losses.items():
        assert v.requires_grad and torch.isfinite(v).all()
        ref = g["losses"][k].item()
        assert abs(v.item() - ref) <= YOLO_TOL[k] * abs(ref) + 1e-3, (k, v.item(), ref)
    sum(losses.values()).backward()
    ps = dict(m.named_parameters())
    e_grad = rel_l2(ps["head.head1.3.bias"].grad, g["grads"]["head.head1.3.bias"])
    print("[zoo] yolov4 head bias grad rel", e_grad)
    assert e_grad < 1.0, e_grad
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
