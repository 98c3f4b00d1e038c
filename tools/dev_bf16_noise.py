"""Dev tool (GPU): how far does ANY bf16 execution of a zoo model land from the fp32 golden logits?

Runs the same module tree three ways on the golden input and prints the rel-L2 distance of the logits to the fixture:
  eager fp32      - torch library kernels in fp32 (must reproduce the fixture: checks the module tree itself)
  eager autocast  - torch library kernels under bf16 autocast (the "natural" bf16 spread of this net at this batch size)
  fused           - this package's CUDA path
"""
import sys

import torch

sys.path.insert(0, ".")
import torch.nn.functional as TF

import holocron_b200 as hb
from holocron_b200.models import _blocks
from holocron_b200.nn import _fused as K

names = sys.argv[1:] or ["darknet53", "cspdarknet53", "darknet24", "darknet19"]
orig_unit = _blocks.conv_bn_act


def eager_unit(x, conv, bn, act, residual=None, res_after_act=False, keep_padded=False):
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None and not res_after_act:
        y = y + residual
    if act is not None:
        y = act(y)
    if residual is not None and res_after_act:
        y = y + residual
    return y


def set_unit(fn):
    for name, mod in list(sys.modules.items()):
        if name.startswith("holocron_b200.models") and hasattr(mod, "conv_bn_act"):
            mod.conv_bn_act = fn


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / b.norm()).item()


golden = torch.load("tests/golden/zoo.pt")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
# conditioning sweep: distance of both bf16 executions to the eager fp32 run on random inputs of several sizes
for name in names:
    for (b, sz) in [(2, 64), (8, 64), (4, 128), (16, 64), (8, 128)]:
        torch.manual_seed(5)
        x = torch.rand(b, 3, sz, sz, device="cuda")
        outs = {}
        for mode in ("eager fp32", "eager autocast", "fused"):
            torch.manual_seed(0)
            m = getattr(hb.models, name)(num_classes=10).cuda().train()
            set_unit(orig_unit if mode == "fused" else eager_unit)
            with torch.no_grad():
                if mode == "eager autocast":
                    with torch.autocast("cuda", dtype=torch.bfloat16):
                        outs[mode] = m(x).float()
                else:
                    outs[mode] = m(x).float()
        set_unit(orig_unit)
        print(f"{name} b{b} {sz}x{sz}: autocast {rel(outs['eager autocast'], outs['eager fp32']):.4f} "
              f"fused {rel(outs['fused'], outs['eager fp32']):.4f}", flush=True)
for name in names:
    g = golden[name]
    x = g["x"].cuda()
    res = {}
    for mode in ("eager fp32", "eager autocast", "fused"):
        torch.manual_seed(0)
        m = getattr(hb.models, name)(num_classes=10).cuda().train()
        set_unit(orig_unit if mode == "fused" else eager_unit)
        try:
            if mode == "eager autocast":
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    out = m(x)
            else:
                out = m(x)
            res[mode] = rel(out, g["logits"])
        except Exception as e:  # noqa: BLE001
            res[mode] = float("nan")
            print("   ", mode, "failed:", repr(e)[:300])
    set_unit(orig_unit)
    print(name, {k: round(v, 5) for k, v in res.items()}, flush=True)
