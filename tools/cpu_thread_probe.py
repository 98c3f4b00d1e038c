"""How many host threads does torch's CPU path want for the oracle train step? (GPU-box host probe)"""
import os, sys, time
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from oracle.models import RepVGGOracle
print("cpu_count", os.cpu_count(), "default threads", torch.get_num_threads())
torch.manual_seed(0)
m = RepVGGOracle("repvgg_a0", num_classes=1000).train()
x = torch.rand(8, 3, 224, 224); t = torch.randint(0, 1000, (8,))
for th in (8, 16, 32, 64):
    if th > (os.cpu_count() or 1):
        break
    torch.set_num_threads(th)
    F.cross_entropy(m(x), t).backward()
    t0 = time.perf_counter()
    F.cross_entropy(m(x), t).backward()
    print(f"threads {th}: {time.perf_counter() - t0:.2f} s / 8 images", flush=True)
