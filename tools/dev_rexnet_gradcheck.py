"""Dev tool (GPU): per-parameter gradient norms of the fused rexnet1_0x vs an eager fp32 execution of the same modules."""
import sys

import torch

sys.path.insert(0, ".")
import torch.nn.functional as TF

import holocron_b200 as hb
from holocron_b200.models.classification import rexnet as R

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
orig_unit, orig_fwd = R.conv_bn_act, R.ReXBlock.forward


def eager_unit(x, conv, bn, act, residual=None, res_after_act=False, keep_padded=False):
    y = conv(x)
    if bn is not None:
        y = bn(y)
    if residual is not None and not res_after_act:
        y = y + residual
    if act is not None:
        y = act(y)
    return y


def eager_block(self, x, keep_padded=False):
    y = x
    for m in self.conv:
        if isinstance(m, R.SEBlock):
            g = y.mean((2, 3), keepdim=True)
            for mm in m.conv:
                g = mm(g)
            y = y * g
        else:
            y = m(y)
    if self.use_shortcut:
        y = torch.cat([y[:, :self.in_channels] + x, y[:, self.in_channels:]], 1)
    return y


g = torch.load("tests/golden/zoo.pt")["rexnet1_0x"]
x, t = g["x"].cuda(), g["t"].cuda()
if len(sys.argv) > 1:
    b = int(sys.argv[1])
    torch.manual_seed(1)
    x = torch.rand(b, 3, 64, 64, device="cuda")
    t = torch.randint(0, 10, (b,), device="cuda")
grads = {}
for mode in ("eager", "fused"):
    torch.manual_seed(0)
    m = hb.models.rexnet1_0x(num_classes=10)
    m.head[0].p = 0.0
    m = m.cuda().train()
    if mode == "eager":
        R.conv_bn_act, R.ReXBlock.forward = eager_unit, eager_block
        from holocron_b200.models import _blocks
        saved = _blocks.conv_bn_act
        _blocks.conv_bn_act = eager_unit
    try:
        out = m(x)
    finally:
        if mode == "eager":
            R.conv_bn_act, R.ReXBlock.forward = orig_unit, orig_fwd
            _blocks.conv_bn_act = saved
    if len(sys.argv) <= 1:
        print(mode, "logits rel to golden", ((out.float().cpu() - g["logits"]).norm() / g["logits"].norm()).item())
    grads["out_" + mode] = out.detach().float()
    TF.cross_entropy(out.float(), t).backward()
    grads[mode] = {k: v.grad.float().clone() for k, v in m.named_parameters()}
print("logits fused vs eager", ((grads["out_fused"] - grads["out_eager"]).norm() / grads["out_eager"].norm()).item())
del grads["out_fused"], grads["out_eager"]
print("golden first-grad norm", g["grads"][g["first"]].norm().item(), "eager", grads["eager"][g["first"]].norm().item(),
      "fused", grads["fused"][g["first"]].norm().item())
for k in grads["eager"]:
    ne, nf = grads["eager"][k].norm().item(), grads["fused"][k].norm().item()
    r = nf / (ne + 1e-30)
    if not 0.5 < r < 2.0 and ne > 1e-5:
        print(f"  {k}: eager {ne:.3e} fused {nf:.3e} ratio {r:.3f}")
