"""Dev tool (GPU): one eager training step of a zoo model between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys

import torch
import torch.nn.functional as TF

sys.path.insert(0, ".")
import holocron_b200 as hb

name = sys.argv[1] if len(sys.argv) > 1 else "rexnet1_0x"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 256
from holocron_b200.distributed import GradBucket
m = getattr(hb.models, name)(num_classes=1000).cuda().to(memory_format=torch.channels_last).train()
bucket = GradBucket(m.parameters())      # gradients added straight into the flat bucket, like bench.py's step
opt = hb.optim.AdaBelief(m.parameters(), lr=1e-3, betas=(0.95, 0.99), eps=1e-6)
x = torch.randn(batch, 3, 224, 224, device="cuda")
t = torch.randint(0, 1000, (batch,), device="cuda")


def step():
    loss = TF.cross_entropy(m(x), t)
    loss.backward()
    opt.step()
    bucket.zero_()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
