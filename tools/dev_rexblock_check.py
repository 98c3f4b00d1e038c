"""Dev tool (GPU): forward / backward of the fused ReXBlock against torch library ops on the same module (fp32)."""
import sys

import torch

sys.path.insert(0, ".")
import torch.nn.functional as TF
from torch import nn

from holocron_b200.models.classification.rexnet import ReXBlock, SEBlock

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def eager(blk, x):
    y = x
    for m in blk.conv:
        if isinstance(m, SEBlock):
            g = y.mean((2, 3), keepdim=True)
            for mm in m.conv:
                g = mm(g)
            y = y * g
        else:
            y = m(y)
    if blk.use_shortcut:
        y = torch.cat([y[:, :blk.in_channels] + x, y[:, blk.in_channels:]], 1)
    return y


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


ok = True
for (cin, ch, t, stride, se, n, hw) in [(16, 16, 1, 1, False, 4, 16), (16, 27, 6, 2, False, 4, 16), (27, 38, 6, 1, False, 4, 16),
                                        (38, 50, 6, 2, True, 4, 16), (50, 61, 6, 1, True, 4, 8), (32, 16, 1, 1, False, 4, 32)]:
    torch.manual_seed(0)
    blk = ReXBlock(cin, ch, t, stride, use_se=se).cuda().train()
    for m in blk.modules():
        if isinstance(m, nn.BatchNorm2d):
            nn.init.uniform_(m.weight, 0.5, 1.5)
            nn.init.normal_(m.bias, 0, 0.2)
    x = torch.randn(n, cin, hw, hw, device="cuda")
    g = None
    res = {}
    for mode in ("eager", "fused"):
        blk.zero_grad()
        xi = x.clone().requires_grad_(True)
        out = eager(blk, xi) if mode == "eager" else blk(xi)
        if g is None:
            g = torch.randn_like(out.float())
        (out.float() * g).sum().backward()
        res[mode] = (out.detach().float(), xi.grad.clone(), {k: v.grad.clone() for k, v in blk.named_parameters()})
    e_out, e_dx = rel(res["fused"][0], res["eager"][0]), rel(res["fused"][1], res["eager"][1])
    worst = max(((rel(res["fused"][2][k], res["eager"][2][k]), k) for k in res["eager"][2]))
    nr = (res["fused"][1].norm() / res["eager"][1].norm()).item()
    good = e_out < 2e-2 and e_dx < 8e-2 and worst[0] < 8e-2
    ok &= good
    print(f"{'OK ' if good else 'BAD'} ReXBlock({cin}->{ch}, t={t}, s={stride}, se={se}) out {e_out:.4f} dx {e_dx:.4f} "
          f"(norm ratio {nr:.3f}) worst param grad {worst[0]:.4f} {worst[1]}", flush=True)
print("ALL OK" if ok else "SOME BAD")
