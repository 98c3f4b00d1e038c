"""Is the end-to-end gradient gap vs the fp32 oracle a bug or bf16 noise? (1) per-block local check at real RepVGG-A0
shapes against an fp32 torch block on the GPU, (2) torch's own bf16-autocast model vs its fp32 self."""
import sys
import copy
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
import holocron_b200 as hb
from holocron_b200.models.classification.repvgg import RepBlock
from oracle.models import RepBlockOracle, RepVGGOracle


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-20)).item()


def block_local(cin, cout, stride, ident, n, h):
    torch.manual_seed(0)
    blk = RepBlock(cin, cout, stride, ident)
    hb.nn.init.init_module(blk)
    ob = RepBlockOracle(cin, cout, stride, ident)
    ob.load_state_dict(blk.state_dict())
    blk, ob = blk.cuda().train(), ob.cuda().train()
    x = torch.randn(n, cin, h, h, device="cuda").relu()
    xb = x.bfloat16()
    xa = xb.clone().requires_grad_(True)
    xo = xb.float().requires_grad_(True)
    y = blk(xa); yo = ob(xo)
    up = torch.randn_like(yo).bfloat16()
    y.backward(up); yo.backward(up.float())
    out = {"y": rel(y, yo), "gx": rel(xa.grad, xo.grad)}
    po = dict(ob.named_parameters())
    for nme, p in blk.named_parameters():
        out[nme] = rel(p.grad, po[nme].grad)
    worst = max(out.values())
    print(f"block {cin}->{cout} s{stride} id{ident} n{n} h{h}: worst {worst:.3e} | " + " ".join(f"{k}={v:.1e}" for k, v in out.items()), flush=True)


for cfg in [(8, 48, 2, False, 8, 224), (48, 48, 1, True, 8, 112), (48, 48, 2, False, 8, 112), (48, 48, 1, True, 8, 56),
            (48, 96, 2, False, 8, 56), (96, 96, 1, True, 8, 28), (96, 192, 2, False, 8, 28), (192, 192, 1, True, 8, 14),
            (192, 1280, 2, False, 8, 14), (1280, 1280, 1, True, 8, 7)]:
    block_local(*cfg)

# (2) torch bf16 autocast vs torch fp32 on the GPU, same weights, batch 8
torch.manual_seed(0)
o32 = RepVGGOracle("repvgg_a0", num_classes=1000).cuda().train()
o16 = copy.deepcopy(o32)
torch.manual_seed(1)
x = torch.rand(8, 3, 224, 224, device="cuda"); t = torch.randint(0, 1000, (8,), device="cuda")
l32 = F.cross_entropy(o32(x), t, label_smoothing=0.1); l32.backward()
with torch.autocast("cuda", dtype=torch.bfloat16):
    l16 = F.cross_entropy(o16(x).float(), t, label_smoothing=0.1)
l16.backward()
print("torch bf16-autocast vs fp32: loss", l16.item(), l32.item())
p32 = dict(o32.named_parameters())
for n in ["head.weight", "features.4.1.branches.0.0.weight", "features.2.2.branches.0.0.weight", "features.0.0.branches.0.0.weight",
          "features.0.0.branches.0.1.weight", "features.3.5.branches.2.bias"]:
    print("   torch-bf16 vs fp32", n, rel(dict(o16.named_parameters())[n].grad, p32[n].grad))
# ours vs fp32 GPU oracle with identical weights
m = hb.models.repvgg_a0(num_classes=1000)
m.load_state_dict(o32.state_dict())
m = m.cuda().train()
lm = F.cross_entropy(m(x), t, label_smoothing=0.1); lm.backward()
print("ours vs fp32: loss", lm.item(), l32.item())
for n in ["head.weight", "features.4.1.branches.0.0.weight", "features.2.2.branches.0.0.weight", "features.0.0.branches.0.0.weight",
          "features.0.0.branches.0.1.weight", "features.3.5.branches.2.bias"]:
    print("   ours vs fp32", n, rel(dict(m.named_parameters())[n].grad, p32[n].grad), " ours vs torch-bf16", rel(dict(m.named_parameters())[n].grad, dict(o16.named_parameters())[n].grad))
