"""Dev diagnostic (GPU): per-parameter gradient agreement of RepVGG-A0 (CUDA path) with the fp32 oracle on the trajectory
test's batch, and the same for torch's own bf16 autocast on the GPU (the 'natural' bf16 spread)."""
import sys
import torch
import torch.nn.functional as TF
sys.path.insert(0, ".")
import holocron_b200 as hb
from oracle.models import RepVGGOracle

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(0)
ours = hb.models.repvgg_a0(num_classes=10)
ref = RepVGGOracle("repvgg_a0", num_classes=10)
ref.load_state_dict(ours.state_dict())
auto = RepVGGOracle("repvgg_a0", num_classes=10)
auto.load_state_dict(ours.state_dict())
g = torch.Generator().manual_seed(21)
x = (torch.rand(16, 3, 64, 64, generator=g) - 0.45) / 0.225
t = torch.randint(0, 10, (16,), generator=g)
ref.train(); TF.cross_entropy(ref(x), t).backward()
auto = auto.cuda().train()
with torch.autocast("cuda", dtype=torch.bfloat16):
    la = TF.cross_entropy(auto(x.cuda()).float(), t.cuda())
la.backward()
ours = ours.cuda().train()
lo = TF.cross_entropy(ours(x.cuda()), t.cuda()); lo.backward()
print("loss ours", lo.item(), "autocast", la.item())
rows = []
for (n, p), (_, q), (_, a) in zip(ours.named_parameters(), ref.named_parameters(), auto.named_parameters()):
    go, gr, ga = p.grad.float().cpu(), q.grad, a.grad.float().cpu()
    rel = lambda u, v: ((u - v).norm() / (v.norm() + 1e-30)).item()
    sign = lambda u, v: (torch.sign(u) == torch.sign(v)).float().mean().item()
    rows.append((n, tuple(p.shape), gr.norm().item(), rel(go, gr), rel(ga, gr), sign(go, gr), sign(ga, gr)))
print(f"{'param':48s} {'|g|':>10s} {'ours':>8s} {'autoc':>8s} {'sgn ours':>8s} {'sgn auto':>8s}")
for r in rows:
    print(f"{r[0]:48s} {r[2]:10.3e} {r[3]:8.4f} {r[4]:8.4f} {r[5]:8.3f} {r[6]:8.3f}")
import numpy as np
print("mean rel ours", np.mean([r[3] for r in rows]), "autocast", np.mean([r[4] for r in rows]))
print("mean sign agreement ours", np.mean([r[5] for r in rows]), "autocast", np.mean([r[6] for r in rows]))
