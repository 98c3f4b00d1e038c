"""Dev diagnostic (GPU): one AdaBelief step of RepVGG-A0 with frozen (calibrated) BatchNorm - CUDA path vs fp32 oracle:
per-parameter gradient error, update error, and the step-2 loss evaluated by the ORACLE on the CUDA path's updated parameters
(separates 'the update differs' from 'the next forward differs')."""
import sys
import torch
import torch.nn.functional as TF
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import holocron_b200 as hb
from oracle.models import RepVGGOracle
from oracle.optim import adabelief_step
import _conditioning as C

torch.manual_seed(0)
ours = hb.models.repvgg_a0(num_classes=10)
ref = RepVGGOracle("repvgg_a0", num_classes=10)
ref.load_state_dict(ours.state_dict())
g = torch.Generator().manual_seed(21)
x = (torch.rand(16, 3, 64, 64, generator=g) - 0.45) / 0.225
t = torch.randint(0, 10, (16,), generator=g)
bns = [m for m in ref.modules() if isinstance(m, torch.nn.BatchNorm2d)]
for m in bns: m.momentum = 1.0
ref.train()
with torch.no_grad(): ref(x)
for m in bns: m.momentum = 0.1
ours.load_state_dict(ref.state_dict())
ref = C.freeze_bn(ref)
lr = 2e-4
p0 = [p.detach().clone() for p in ref.parameters()]
l1 = TF.cross_entropy(ref(x), t); l1.backward()
gref = [p.grad.clone() for p in ref.parameters()]
for p in ref.parameters():
    adabelief_step(p.data, p.grad, torch.zeros_like(p), torch.zeros_like(p), 1, lr, 0.95, 0.99, 1e-6); p.grad = None
with torch.no_grad(): l2 = TF.cross_entropy(ref(x), t)
ours = C.freeze_bn(ours.cuda())
opt = hb.optim.AdaBelief(ours.parameters(), lr=lr, betas=(0.95, 0.99), eps=1e-6)
o1 = TF.cross_entropy(ours(x.cuda()), t.cuda()); o1.backward()
gours = [None if p.grad is None else p.grad.detach().float().cpu().clone() for p in ours.parameters()]
opt.step(); opt.zero_grad()
with torch.no_grad(): o2 = TF.cross_entropy(ours(x.cuda()), t.cuda())
print("step1 loss oracle", l1.item(), "ours", o1.item(), "| step2 oracle", l2.item(), "ours", o2.item())
names = [n for n, _ in ours.named_parameters()]
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
bad = 0
for n, go, gr, p, q, q0 in zip(names, gours, gref, ours.parameters(), ref.parameters(), p0):
    if go is None:
        print("NO GRAD", n); bad += 1; continue
    du, dr = p.detach().float().cpu() - q0, q.detach() - q0
    sign = (torch.sign(go) == torch.sign(gr)).float().mean().item()
    if rel(go, gr) > 0.5 or rel(du, dr) > 0.7:
        print(f"{n:44s} grad rel {rel(go, gr):.3f} sign {sign:.3f} update rel {rel(du, dr):.3f} |dr| {dr.abs().mean():.2e} |du| {du.abs().mean():.2e}")
# oracle forward on OUR updated parameters
ref2 = RepVGGOracle("repvgg_a0", num_classes=10)
sd = {k: v.detach().float().cpu() for k, v in ours.state_dict().items()}
ref2.load_state_dict(sd); ref2 = C.freeze_bn(ref2)
with torch.no_grad(): print("oracle forward on CUDA-updated params:", TF.cross_entropy(ref2(x), t).item())
import numpy as np
print("mean grad rel", np.mean([rel(a, b) for a, b in zip(gours, gref) if a is not None]), "missing", bad)
