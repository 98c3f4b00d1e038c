"""GPU parity tests (through the C ABI) for activations, losses and box operators: CUDA path vs the golden fixtures
produced by the unmodified reference and vs the CPU oracle on seeded inputs.

Tolerances: fp32 in/out -> rtol 1e-5 (well inside north_star's 1e-3); exact `==` for the reference's own box vectors;
bf16 I/O -> one bf16 ulp (2^-8 relative) on the output."""
import math

import pytest
import torch

import holocron_b200 as hb
from holocron_b200.nn import functional as F
from holocron_b200.ops import boxes as B
from oracle import boxes as OB
from oracle import functional as OF

from conftest import load_golden

pytestmark = pytest.mark.gpu


def close(a, b, rtol=1e-5, atol=1e-6):
    torch.testing.assert_close(a.detach().cpu().float(), b.detach().cpu().float(), rtol=rtol, atol=atol)


def grad_of(fn, *inputs):
    ins = [t.clone().cuda().requires_grad_(True) for t in inputs]
    out = fn(*ins)
    gs = torch.autograd.grad(out.sum(), ins, allow_unused=True)
    return out.detach(), gs


def test_activations_vs_golden():
    g = load_golden("activations")
    y, (gx,) = grad_of(lambda t: F.hard_mish(t), g["x"])
    close(y, g["hard_mish"]); close(gx, g["hard_mish_grad"])
    for beta in (1.0, 0.5):
        y, (gx,) = grad_of(lambda t: F.nl_relu(t, beta=beta), g["x"])
        close(y, g[f"nl_relu_b{beta}"]); close(gx, g[f"nl_relu_b{beta}_grad"])


def test_activation_inplace_aliasing_and_grads():
    # reference tests/test_nn_activation.py:26-27: inplace returns a tensor aliasing the input
    for fn in (F.hard_mish, F.nl_relu):
        x = torch.rand(4, 3, 32, 32, device="cuda")
        ref = fn(x.clone())
        out = fn(x, inplace=True)
        assert out.data_ptr() == x.data_ptr() and out.shape == x.shape
        close(out, ref, 0, 0)
    # in-place under autograd on a non-leaf gives the same gradient as out-of-place
    g = load_golden("activations")
    for fn, key in ((F.hard_mish, "hard_mish_grad"), (F.nl_relu, "nl_relu_b1.0_grad")):
        x = g["x"].cuda().requires_grad_(True)
        y = fn(x * 1.0, inplace=True)
        y.sum().backward()
        close(x.grad, g[key], 1e-5, 1e-6)
    assert repr(hb.nn.HardMish()) == "HardMish()" and repr(hb.nn.NLReLU(inplace=True)) == "NLReLU(inplace=True)"


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_activations_low_precision_and_odd_sizes(dtype):
    torch.manual_seed(0)
    for n in (1, 7, 4097, 1 << 20):
        x = (torch.randn(n) * 3).to(dtype)
        for fn, ofn in ((F.hard_mish, OF.hard_mish), (F.nl_relu, OF.nl_relu)):
            y = fn(x.cuda())
            assert y.dtype == dtype
            ref = ofn(x.float())
            close(y, ref, 2 ** -7 if dtype == torch.bfloat16 else 2 ** -9, 1e-2)
    # unaligned view (scalar path) must match the aligned result
    base = torch.randn(1025, device="cuda")
    close(F.hard_mish(base[1:]), F.hard_mish(base[1:].clone()), 0, 0)


def test_activation_full_size_property():
    # Darknet-sized activation (64 x 64 x 112 x 112, bf16). Size-independent properties:
    # hard_mish(x) == x for x >= 0 ... no: == x * min(1, (x+2)/2); == 0 for x <= -2; nl_relu(x) == 0 for x <= 0.
    x = torch.randn(64, 64, 112, 112, device="cuda", dtype=torch.bfloat16) * 2
    y = F.hard_mish(x)
    assert torch.all(y[x <= -2] == 0) and torch.all(y[x >= 0] == x[x >= 0])
    z = F.nl_relu(x)
    assert torch.all(z[x <= 0] == 0) and torch.all(z[x > 0] > 0)
    idx = torch.randint(0, x.numel(), (4096,), device="cuda")
    close(y.view(-1)[idx], OF.hard_mish(x.view(-1)[idx].float().cpu()), 2 ** -7, 1e-2)


def test_cpu_tensors_are_rejected():
    with pytest.raises(hb.HolocronB200Error):
        F.hard_mish(torch.randn(4))
    with pytest.raises(hb.HolocronB200Error):
        B.diou_loss(torch.rand(2, 4), torch.rand(2, 4))


# ------------------------------------------------------------------------------------------------ losses
def test_losses_vs_golden():
    g = load_golden("losses")
    for tag in ("cls", "seg"):
        x, t, w = g[f"{tag}_x"], g[f"{tag}_t"].cuda(), g[f"{tag}_w"].cuda()
        for red in ("mean", "sum", "none"):
            for ii in (-100, 1):
                for use_w in (False, True):
                    key = f"{tag}_{red}_ii{ii}_w{int(use_w)}"
                    wt = w if use_w else None
                    y, (gx,) = grad_of(lambda a: F.focal_loss(a, t, wt, ii, red, 2.0), x)
                    assert y.shape == g["focal_" + key].shape
                    close(y, g["focal_" + key], 2e-5, 1e-6); close(gx, g["focal_grad_" + key], 1e-4, 1e-6)
                    y, (gx,) = grad_of(lambda a: F.poly_loss(a, t, 2.0, wt, ii, red), x)
                    assert y.shape == g["poly_" + key].shape
                    close(y, g["poly_" + key], 2e-5, 1e-6); close(gx, g["poly_grad_" + key], 1e-4, 1e-6)
        y, (gx,) = grad_of(lambda a: F.focal_loss(a, t, None, -100, "mean", 0.5), x)
        close(y, g[f"focal_{tag}_gamma0.5"], 2e-5, 1e-6); close(gx, g[f"focal_grad_{tag}_gamma0.5"], 1e-4, 1e-6)
        soft = g[f"{tag}_soft"].cuda()
        for red in ("mean", "sum", "none"):
            for ii in (-100, 1):
                y, (gx,) = grad_of(lambda a: F.poly_loss(a, soft, 2.0, None, ii, red), x)
                assert y.shape == g[f"polysoft_{tag}_{red}_ii{ii}"].shape
                close(y, g[f"polysoft_{tag}_{red}_ii{ii}"], 2e-5, 1e-6)
                close(gx, g[f"polysoft_grad_{tag}_{red}_ii{ii}"], 1e-4, 1e-6)
    y, (gx,) = grad_of(lambda a: F.poly_loss(a, g["cls_soft"].cuda(), 1.5, g["cls_w"].cuda(), -100, "mean"), g["cls_x"])
    close(y, g["polysoft_cls_w"], 2e-5, 1e-6); close(gx, g["polysoft_grad_cls_w"], 1e-4, 1e-6)
    for gamma in (1.0, 2.0):
        for use_w in (False, True):
            wt = g["seg_w"].cuda() if use_w else None
            y, (gx,) = grad_of(lambda a: F.dice_loss(a, g["seg_onehot"].cuda(), wt, gamma), g["seg_prob"])
            close(y, g[f"dice_seg_g{gamma}_w{int(use_w)}"], 2e-5, 1e-6)
            close(gx, g[f"dice_grad_seg_g{gamma}_w{int(use_w)}"], 1e-4, 1e-7)


def _loss_harness(fn, same_loss=0.0, multi_label=False):
    """The reference's shared loss harness (tests/test_nn_loss.py:9-47) run against the CUDA implementation."""
    num_batches, num_classes = 2, 4
    x = torch.ones(num_batches, num_classes, device="cuda")
    x[:, 0, ...] = 100
    x.requires_grad_(True)
    if multi_label:
        target = torch.zeros_like(x)
        target[:, 0] = 1.0
    else:
        target = torch.zeros(num_batches, dtype=torch.long, device="cuda")
    assert abs(fn(x, target).item() - same_loss) < 1e-3
    assert torch.allclose(fn(x, target, reduction="none"), same_loss * torch.ones(num_batches, dtype=x.dtype, device="cuda"),
                          atol=1e-3)
    x = torch.rand(num_batches, num_classes, device="cuda", requires_grad=True)
    target = torch.rand(x.shape, device="cuda") if multi_label else (num_classes * torch.rand(num_batches, device="cuda")).to(torch.long)
    weights = torch.ones(num_classes, device="cuda")
    assert fn(x, target).item() == fn(x, target, weight=weights).item()
    assert fn(x, target).item() == fn(x, target, ignore_index=num_classes).item()
    ignore_index = torch.unique(target.argmax(dim=1))[0].item() if multi_label else torch.unique(target)[0].item()
    assert fn(x, target).item() != fn(x, target, ignore_index=ignore_index).item()
    loss = fn(x, target, ignore_index=0)
    loss.backward()
    assert torch.allclose(fn(x, target, reduction="sum"), fn(x, target, reduction="none").sum(), atol=1e-6)
    assert torch.allclose(fn(x, target, reduction="mean"), fn(x, target, reduction="sum") / target.shape[0], atol=1e-6)


def test_loss_reference_harness():
    _loss_harness(F.focal_loss)
    _loss_harness(F.poly_loss)
    _loss_harness(F.poly_loss, multi_label=True)
    torch.manual_seed(0)
    x = torch.rand(2, 4, 20, 20, device="cuda")
    t = torch.randint(0, 4, (2, 20, 20), device="cuda")
    close(F.focal_loss(x, t, gamma=0.0), torch.nn.functional.cross_entropy(x, t), 1e-5, 1e-6)
    ones = torch.ones(2, 4, 20, 20, device="cuda")  # equal probabilities (reference test_nn_loss.py:62-65)
    close((1 - 1 / 4) * F.focal_loss(ones, t, gamma=0), F.focal_loss(ones, t, gamma=1), 1e-5, 1e-6)
    xg = torch.rand(2, 4, 20, 20, device="cuda", requires_grad=True)
    F.dice_loss(xg, torch.rand(2, 4, 20, 20, device="cuda")).backward()
    cw = torch.ones(4, device="cuda"); cw[0] = 2
    F.dice_loss(xg, torch.rand(2, 4, 20, 20, device="cuda"), weight=cw).backward()
    F.poly_loss(xg, t, weight=cw).backward()
    with pytest.raises(TypeError):
        F.poly_loss(x, t.int())
    with pytest.raises(ValueError):
        F.poly_loss(x, torch.rand(2, 5, 20, 20, device="cuda"))


def test_losses_segmentation_size_vs_oracle():
    # 16 x 21 x 256 x 256 segmentation logits (bf16). Properties at full size + oracle on a slice.
    torch.manual_seed(1)
    x = torch.randn(16, 21, 256, 256, device="cuda", dtype=torch.bfloat16)
    t = torch.randint(0, 21, (16, 256, 256), device="cuda")
    none = F.focal_loss(x, t, reduction="none").float()
    close(F.focal_loss(x, t, reduction="sum").float(), none.sum(), 5e-3, 0)
    close(F.focal_loss(x, t, reduction="mean").float(), none.mean(), 5e-3, 0)
    ref = OF.focal_loss(x[:2].float().cpu(), t[:2].cpu(), reduction="none")
    close(none[:2], ref, 2 ** -7, 1e-3)
    p = torch.softmax(x.float(), 1)
    oh = torch.nn.functional.one_hot(t, 21).movedim(-1, 1).float()
    d = F.dice_loss(p, oh)
    close(d, OF.dice_loss(p.cpu(), oh.cpu()), 1e-4, 1e-6)


# ------------------------------------------------------------------------------------------------ boxes
KAT = torch.tensor([[0, 0, 100, 100], [50, 50, 100, 100], [50, 50, 150, 150], [100, 100, 200, 200]], dtype=torch.float32)


def test_box_known_answers_exact():
    """The reference's own exact-value test vectors (tests/test_ops.py:16-76), `==` on fp32."""
    b = KAT.cuda()
    pen = B.iou_penalty(b, b)
    assert pen.shape == (4, 4)
    assert all(pen[i, i].item() == 0 for i in range(4))
    assert pen[0, 1].item() == 25**2 / 100**2 and pen[0, 3].item() == 100**2 / 200**2
    assert pen[0, 2].item() == pen[2, 3].item()
    diou = B.diou_loss(b, b)
    assert all(diou[i, i].item() == 0.0 for i in range(4))
    assert diou[0, 1].item() == 1 - 0.25 + 25**2 / 100**2 and diou[0, 3].item() == 1 + 100**2 / 200**2
    assert diou[0, 2].item() == diou[2, 3].item()
    giou = B.box_giou(b, b)
    assert all(giou[i, i].item() == 1.0 for i in range(4))
    assert giou[0, 1].item() == 0.25 and giou[0, 3].item() == -(200**2 - 2 * 100**2) / 200**2
    assert giou[0, 2].item() == giou[2, 3].item()
    close(B.aspect_ratio(b), math.pi / 4 * torch.ones(4), 1e-6, 0)
    assert torch.equal(B.aspect_ratio_consistency(b, b).cpu(), torch.zeros(4, 4))
    ciou = B.ciou_loss(b, b)
    assert all(ciou[i, i].item() == 0.0 for i in range(4)) and ciou[0, 2].item() == ciou[2, 3].item()
    with pytest.raises(AssertionError):
        B.box_giou(torch.tensor([[10.0, 10, 0, 0]], device="cuda"), b)


def test_boxes_vs_golden_bit_exact_and_grads():
    g = load_golden("boxes")
    b1, b2 = g["b1"].cuda(), g["b2"].cuda()
    for name, (a, b) in (("kat", (g["kat"].cuda(), g["kat"].cuda())), ("rnd", (b1, b2))):
        assert torch.equal(B.box_giou(a, b).cpu(), g[f"{name}_giou"])
        assert torch.equal(B.iou_penalty(a, b).cpu(), g[f"{name}_penalty"])
        assert torch.equal(B.diou_loss(a, b).cpu(), g[f"{name}_diou"])
        assert torch.equal(B.ciou_loss(a, b).cpu(), g[f"{name}_ciou"])
        close(B.aspect_ratio_consistency(a, b), g[f"{name}_arc"], 1e-5, 1e-7)
    for fn in ("box_giou", "diou_loss", "ciou_loss"):
        _, (g1, g2) = grad_of(getattr(B, fn), g["b1"], g["b2"])
        close(g1, g[f"rnd_{fn}_grad1"], 1e-4, 1e-6); close(g2, g[f"rnd_{fn}_grad2"], 1e-4, 1e-6)
    a = g["b1"].cuda().requires_grad_(True); b = g["b2"].cuda().requires_grad_(True)
    (B.ciou_loss(a, b) * g["up"].cuda()).sum().backward()
    close(a.grad, g["rnd_ciou_wgrad1"], 1e-4, 1e-6); close(b.grad, g["rnd_ciou_wgrad2"], 1e-4, 1e-6)
    # bf16 inputs -> fp32 output like the reference; empty and ragged shapes
    assert B.diou_loss(b1.bfloat16(), b2.bfloat16()).dtype == torch.float32
    assert B.diou_loss(b1[:0], b2).shape == (0, 23) and B.diou_loss(b1, b2[:0]).shape == (37, 0)
    assert B.diou_loss(b1[:1], b2[:1]).shape == (1, 1)


def test_boxes_large_vs_oracle():
    torch.manual_seed(3)
    xy = torch.rand(2000, 2); wh = torch.rand(2000, 2) * 0.2 + 0.01
    b1 = torch.cat([xy, xy + wh], 1)
    xy = torch.rand(50, 2); wh = torch.rand(50, 2) * 0.2 + 0.01
    b2 = torch.cat([xy, xy + wh], 1)
    assert torch.equal(B.diou_loss(b1.cuda(), b2.cuda()).cpu(), OB.diou_loss(b1, b2))
    assert torch.equal(B.box_giou(b1.cuda(), b2.cuda()).cpu(), OB.box_giou(b1, b2))
    # YOLOv4-style use (yolov4.py:401-403): min over GT then sum, gradient to the predictions
    a = b1.cuda().requires_grad_(True)
    B.ciou_loss(a, b2.cuda()).min(dim=1).values.sum().backward()
    ao = b1.clone().requires_grad_(True)
    OB.ciou_loss(ao, b2).min(dim=1).values.sum().backward()
    close(a.grad, ao.grad, 1e-3, 1e-5)


@pytest.mark.parametrize("m,n", [(33, 64), (5, 3), (1, 1), (17, 516), (40, 1028), (3, 7)])
def test_boxes_tile_edges_bit_exact(m, n):
    """Tiled kernel: 4 columns per thread (128-bit store when N % 4 == 0, scalar tail otherwise), 16 rows per block, 512
    columns per block in x - every edge, all five operators, bit for bit against the CPU oracle."""
    torch.manual_seed(m * 1000 + n)
    xy = torch.rand(m, 2); wh = torch.rand(m, 2) * 0.3 + 0.01
    b1 = torch.cat([xy, xy + wh], 1)
    xy = torch.rand(n, 2); wh = torch.rand(n, 2) * 0.3 + 0.01
    b2 = torch.cat([xy, xy + wh], 1)
    assert torch.equal(B.box_iou(b1.cuda(), b2.cuda()).cpu(), OB.box_iou(b1, b2))
    assert torch.equal(B.box_giou(b1.cuda(), b2.cuda()).cpu(), OB.box_giou(b1, b2))
    assert torch.equal(B.diou_loss(b1.cuda(), b2.cuda()).cpu(), OB.diou_loss(b1, b2))
    assert torch.equal(B.ciou_loss(b1.cuda(), b2.cuda()).cpu(), OB.ciou_loss(b1, b2))
    # a view whose storage offset is not a multiple of 16 bytes is re-aligned by the wrapper
    flat = torch.cat([torch.zeros(1), b1.flatten()]).cuda()
    assert torch.equal(B.diou_loss(flat[1:].view(m, 4), b2.cuda()).cpu(), OB.diou_loss(b1, b2))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k,hw", [(3, (8, 12)), (12, (4, 8)), (21, (16, 8)), (30, (6, 4)), (40, (4, 4)), (21, (5, 7))])
def test_losses_register_resident_path_vs_oracle(dtype, k, hw):
    """S > 1 and K <= 32 take the register-resident kernels (KMAX 8 / 16 / 24 / 32; 8-byte vectors: S % 2 == 0 for fp32,
    S % 4 == 0 for bf16), everything else the one-position-per-thread kernels: both against the CPU oracle on the SAME
    (dtype-rounded) logits, forward and backward, with class weights and an ignored class."""
    torch.manual_seed(k)
    n = 3
    x = (torch.randn(n, k, *hw) * 2).to(dtype)
    t = torch.randint(0, k, (n, *hw))
    w = torch.rand(k) + 0.5
    soft = torch.softmax(torch.randn(n, k, *hw), 1).to(dtype)
    tol = dict(rtol=2e-5, atol=2e-6) if dtype == torch.float32 else dict(rtol=2 ** -7, atol=2e-3)
    gtol = dict(rtol=1e-4, atol=1e-6) if dtype == torch.float32 else dict(rtol=2 ** -6, atol=2e-3)
    cases = [
        (lambda a, tt: F.focal_loss(a, tt, w.to(a.device), 1, "mean", 2.0), lambda a: OF.focal_loss(a, t, w, 1, "mean", 2.0), t),
        (lambda a, tt: F.focal_loss(a, tt, None, -100, "none", 0.5), lambda a: OF.focal_loss(a, t, None, -100, "none", 0.5), t),
        (lambda a, tt: F.poly_loss(a, tt, 2.0, w.to(a.device), 1, "sum"), lambda a: OF.poly_loss(a, t, 2.0, w, 1, "sum"), t),
        (lambda a, tt: F.poly_loss(a, tt, 1.5, None, 1, "mean"), lambda a: OF.poly_loss(a, soft.float(), 1.5, None, 1, "mean"), soft),
        (lambda a, tt: F.poly_loss(a, tt, 2.0, None, -100, "none"), lambda a: OF.poly_loss(a, soft.float(), 2.0, None, -100, "none"), soft),
    ]
    for ours, ref, tgt in cases:
        a = x.clone().cuda().requires_grad_(True)
        y = ours(a, tgt.cuda())
        up = torch.rand(y.shape, generator=torch.Generator().manual_seed(1)) + 0.5
        (y.float() * up.cuda()).sum().backward()
        ao = x.float().clone().requires_grad_(True)
        yo = ref(ao)
        (yo * up).sum().backward()
        torch.testing.assert_close(y.detach().float().cpu(), yo.detach(), **tol)
        torch.testing.assert_close(a.grad.float().cpu(), ao.grad, **gtol)


def test_dice_is_deterministic_and_matches_oracle_on_ragged_planes():
    """Per-block partial sums folded in a fixed order: two runs are bit-identical (the first version added doubles
    atomically); vector (S % 8 == 0 bf16, S % 4 == 0 fp32) and scalar planes against the oracle."""
    torch.manual_seed(5)
    for dtype, shape in ((torch.float32, (4, 21, 64, 64)), (torch.float32, (3, 5, 7, 9)), (torch.bfloat16, (4, 21, 32, 40)),
                         (torch.bfloat16, (2, 3, 5, 5))):
        p = torch.softmax(torch.randn(*shape), 1).to(dtype)
        oh = torch.nn.functional.one_hot(torch.randint(0, shape[1], (shape[0], *shape[2:])), shape[1]).movedim(-1, 1).to(dtype)
        w = torch.rand(shape[1]) + 0.5
        runs = []
        for _ in range(2):
            a = p.clone().cuda().requires_grad_(True)
            y = F.dice_loss(a, oh.cuda(), w.cuda(), 2.0)
            y.backward()
            runs.append((y.detach().clone(), a.grad.clone()))
        assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])
        ao = p.float().clone().requires_grad_(True)
        yo = OF.dice_loss(ao, oh.float(), w, 2.0)
        yo.backward()
        fp32 = dtype == torch.float32
        close(runs[0][0], yo, 2e-5 if fp32 else 2 ** -7, 1e-6 if fp32 else 1e-3)
        close(runs[0][1], ao.grad, 1e-4 if fp32 else 2 ** -6, 1e-8 if fp32 else 1e-6)
