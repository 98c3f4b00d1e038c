"""Shared fixture conditioning for the model-zoo parity tests (used by tests/golden/make_golden.py --zoo on the reference's
models and by the tests on this package's models: identical module trees -> identical parameter walk).

A freshly initialised network has BatchNorm weight 1 / bias 0 / running statistics (0, 1) and - for YOLOv4 - all-zero
output convolutions, which would make several code paths invisible to a parity check (affine folding, eval-mode
statistics, every gradient behind a zero filter). `condition` gives every BatchNorm layer distinct affine parameters and
running statistics and the YOLOv4 output convolutions small non-zero filters, all from one seeded CPU generator."""
import torch
from torch import nn


def condition(model: nn.Module, seed: int = 1234, zero_convs: bool = True) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for mod in model.modules():
            if isinstance(mod, nn.BatchNorm2d):
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.75)
                mod.bias.copy_(torch.rand(mod.bias.shape, generator=g) * 0.4 - 0.2)
                mod.running_mean.copy_(torch.rand(mod.running_mean.shape, generator=g) * 0.4 - 0.2)
                mod.running_var.copy_(torch.rand(mod.running_var.shape, generator=g) * 1.0 + 0.5)
            elif type(mod).__name__ == "LayerScale":
                # ConvNeXt: a block's branch is scaled by 1e-6 at initialisation - invisible to any output comparison
                mod.weight.copy_(torch.rand(mod.weight.shape, generator=g) * 0.5 + 0.25)
            elif zero_convs and isinstance(mod, nn.Conv2d) and mod.bias is not None and float(mod.weight.abs().sum()) == 0.0:
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) * 0.02)
                mod.bias.copy_(torch.rand(mod.bias.shape, generator=g) * 0.2 - 0.1)
    return model


def freeze_bn(model: nn.Module) -> nn.Module:
    """Training-mode model whose BatchNorm layers use their running statistics (the reference's trainer does this with
    holocron/trainer/utils.py freeze_bn); gradients still flow to every parameter."""
    model.train()
    for mod in model.modules():
        if isinstance(mod, nn.BatchNorm2d):
            mod.eval()
    return model


CLS = {
    # name: (eval batch, train batch, image size)
    "darknet24": (4, 8, 64), "darknet19": (4, 8, 64), "darknet53": (4, 8, 64), "cspdarknet53": (4, 8, 64),
    "cspdarknet53_mish": (4, 8, 64), "rexnet1_0x": (8, 8, 64), "repvgg_a0": (4, 8, 64),
}


# the ResNet family - basic blocks, bottlenecks with projection shortcuts, the ResNet-D stem and
# average-pooled shortcuts, ResNeXt's grouped 3x3 units. Separate fixture file (tests/golden/zoo_resnet.pt).
CLS_RESNET = {"resnet18": (4, 8, 64), "resnet50d": (4, 8, 64), "resnext50_32x4d": (4, 8, 64),
              "mobileone_s0": (4, 8, 64)}     # MobileOne: re-parametrisable depth-wise / point-wise branch sums (up to 6 branches)


# SURVEY §8 f3, continued: Res2Net (hierarchical 26-channel slices), SKNet (selective-kernel units: grouped dilated paths +
# soft attention), ConvNeXt (depth-wise 7x7, LayerNorm, GELU, LayerScale, patchify convolutions). tests/golden/zoo_f3.pt.
CLS_F3 = {"res2net50_26w_4s": (4, 8, 64), "sknet50": (4, 8, 64), "convnext_atto": (4, 8, 64)}
# ... and the rest of reference models/classification/*.py: TridentNet (one filter shared by three dilation branches stacked on
# the channel axis), PyConvResNet (pyramidal 3x3..9x9 grouped convolutions). tests/golden/zoo_f3b.pt.
CLS_F3B = {"tridentnet50": (4, 8, 64), "pyconv_resnet50": (4, 8, 64), "pyconvhg_resnet50": (4, 8, 64)}


# U-Net family (reference models/segmentation/unet.py, unetpp.py): plain U-Net, the nested UNet+ / UNet++ grids, DynamicUNet on its
# own contracting path and on this package's ReXNet-1.3x (odd channel counts at every tap). tests/golden/zoo_seg.pt.
SEG = ("unet", "unetp", "unetpp", "unet2", "unet_rexnet13")


def seg_inputs():
    g = torch.Generator().manual_seed(16)
    x = torch.rand(2, 3, 64, 64, generator=g)
    mask = torch.randint(0, 5, (2, 64, 64), generator=g)
    return x, mask


def seg_kwargs(name: str):
    return dict(num_classes=5, **({"pretrained_backbone": False} if "rexnet" in name else {}))


def cls_inputs(name: str, mode: str):
    be, bt, size = {**CLS, **CLS_RESNET, **CLS_F3, **CLS_F3B}[name]
    b = be if mode == "eval" else bt
    g = torch.Generator().manual_seed(11 if mode == "eval" else 12)
    x = (torch.rand(b, 3, size, size, generator=g) - 0.45) / 0.225
    t = torch.randint(0, 10, (b,), generator=g)
    return x, t


def yolo_inputs():
    g = torch.Generator().manual_seed(13)
    x = torch.rand(2, 3, 128, 128, generator=g)
    # three well-separated boxes per image (different cells at every scale, distinct shapes -> unambiguous anchors)
    centres = torch.tensor([[0.2, 0.25], [0.55, 0.6], [0.8, 0.3]])
    target = []
    for i in range(2):
        wh = torch.tensor([[0.08, 0.12], [0.22, 0.3], [0.45, 0.4]]) * (1.0 + 0.1 * i)
        c = centres + 0.03 * i
        boxes = torch.cat([c - wh / 2, c + wh / 2], 1).clamp(0, 1)
        target.append({"boxes": boxes, "labels": torch.tensor([3 + i, 17, 42])})
    return x, target


def yolo_dup_inputs():
    """Like yolo_inputs, but in image 0 two boxes share their cell AND their best-shape anchor at every scale (same centre,
    nearly the same shape): the reference's boolean masks collapse them into ONE assigned slot; and image 1 has a single box."""
    x, _ = yolo_inputs()
    c = torch.tensor([[0.30, 0.40], [0.30, 0.40], [0.70, 0.65]])
    wh = torch.tensor([[0.20, 0.26], [0.21, 0.25], [0.10, 0.12]])
    t0 = {"boxes": torch.cat([c - wh / 2, c + wh / 2], 1), "labels": torch.tensor([5, 9, 33])}
    c1, wh1 = torch.tensor([[0.5, 0.5]]), torch.tensor([[0.4, 0.3]])
    t1 = {"boxes": torch.cat([c1 - wh1 / 2, c1 + wh1 / 2], 1), "labels": torch.tensor([61])}
    return x, [t0, t1]


def yolo12_inputs(name: str):
    """YOLOv1 (448 x 448: its classifier is sized for the 7 x 7 grid) / YOLOv2 (128 x 128 -> 4 x 4 grid) inputs: the
    shared-slot targets of yolo_dup_inputs with VOC-range labels (two boxes of image 0 fall into the same cell)."""
    g = torch.Generator().manual_seed(15)
    size = 448 if name == "yolov1" else 128
    x = torch.rand(2, 3, size, size, generator=g)
    _, target = yolo_dup_inputs()
    return x, [{"boxes": t["boxes"], "labels": t["labels"] % 20} for t in target]


def unet_inputs():
    g = torch.Generator().manual_seed(14)
    x = torch.rand(2, 3, 64, 64, generator=g)
    mask = torch.randint(0, 21, (2, 64, 64), generator=g)
    return x, mask


# Training-mode probes: an early container module (same path in the reference's and in this package's tree) whose output is
# compared tightly. Batch-statistics BatchNorm removes the per-channel mean of every convolution output - a component
# that carries signal energy but no rounding noise - so the relative bf16 rounding noise of a random-init network grows
# by ~1.2x per layer (measured: tools/dev_fixture_conditioning.py, and two fp32-arithmetic executions that share every bf16
# storage point but differ by 1e-6 in summation order already land 0.05-0.17 apart on the 19..53-layer nets). End to end
# the training-mode check is therefore: probe <= 2e-2 after 4-7 layers, loss <= 5e-2 at full depth, plus the per-launch
# teacher-forcing checks at every depth; the frozen-BatchNorm ("eval") fixtures carry the tight full-depth comparison.
PROBE = {
    "darknet24": "features.layers.0", "darknet19": "features.layers.0", "darknet53": "features.layers.0",
    "cspdarknet53": "features.stages.0", "cspdarknet53_mish": "features.stages.0", "rexnet1_0x": "features.4",
    "repvgg_a0": "features.1", "unet3p": "encoder.1", "yolov4": "backbone.stages.0",
    "unet": "encoder.1", "unetp": "encoder.1", "unetpp": "encoder.1", "unet2": "encoder.1", "unet_rexnet13": "encoder.4",
    "yolov1": "backbone.layers.0", "yolov2": "backbone.layers.1",
    "tridentnet50": "features.5.0", "pyconv_resnet50": "features.3.0", "pyconvhg_resnet50": "features.3.0",
    "res2net50_26w_4s": "features.4.0", "sknet50": "features.4.0", "convnext_atto": "features.2.0",
    "resnet18": "features.4", "resnet50d": "features.10.0", "resnext50_32x4d": "features.4.0", "mobileone_s0": "features.1.0",
}


def capture(model: nn.Module, path: str, store: dict):
    """Forward hook storing the (detached) output of sub-module ``path`` in ``store['probe']``."""
    mod = model.get_submodule(path)

    def hook(_m, _inp, out):
        store["probe"] = (out[0] if isinstance(out, (tuple, list)) else out).detach()
    return mod.register_forward_hook(hook)


def head_rows(t: torch.Tensor, rows: int = 8) -> torch.Tensor:
    """Fixture diet for the YOLOv1 / YOLOv2 gradients (1024 x 1024 x 3 x 3 filters, a 1470 x 512 classifier): the first
    ``rows`` output rows of a large tensor stand for the whole (generator and tests apply the same cut)."""
    return t[:rows] if t.numel() > 100_000 else t
