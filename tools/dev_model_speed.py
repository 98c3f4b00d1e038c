"""Dev tool (GPU): training-step throughput of the other zoo models (eager launches, AdaBelief, synthetic data)."""
import sys
import time

import torch
import torch.nn.functional as TF

sys.path.insert(0, ".")
import holocron_b200 as hb

cfgs = [("rexnet1_0x", 128, 224), ("darknet53", 64, 224), ("cspdarknet53", 64, 224), ("darknet19", 128, 224), ("repvgg_a1", 256, 224)]
if len(sys.argv) > 1:
    cfgs = [c for c in cfgs if c[0] in sys.argv[1:]]
for name, batch, size in cfgs:
    torch.manual_seed(0)
    m = getattr(hb.models, name)(num_classes=1000).cuda().to(memory_format=torch.channels_last).train()
    opt = hb.optim.AdaBelief(m.parameters(), lr=1e-3)
    x = torch.randn(batch, 3, size, size, device="cuda")
    t = torch.randint(0, 1000, (batch,), device="cuda")

    def step():
        loss = TF.cross_entropy(m(x), t)
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        loss = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    # torch eager reference on the same modules? not available for fused trees; report absolute numbers
    print(f"{name}: batch {batch} {size}x{size}: {dt*1e3:.1f} ms/step {batch/dt:.0f} img/s loss {loss.item():.3f} "
          f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB", flush=True)
    del m, opt, x, t
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats()
