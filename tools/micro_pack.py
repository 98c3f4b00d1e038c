"""Times the multi-tensor filter packing launch (hb_pack_conv_weights_multi) of a model: the first kernel of every training step."""
import sys
import torch
sys.path.insert(0, ".")
import holocron_b200 as hb
from holocron_b200.nn import _fused as K

name = sys.argv[1] if len(sys.argv) > 1 else "repvgg_a0"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 32
torch.manual_seed(0)
m = getattr(hb.models, name)(num_classes=1000).cuda().to(memory_format=torch.channels_last).train()
x = torch.randn(batch, 3, 224, 224, device="cuda")
for _ in range(2):       # registers every filter in the pack table
    m(x).float().sum().backward()
    torch.autograd.graph.increment_version(list(m.parameters()))
assert K._pack_table.repack_all()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(42)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ts = []
for i in range(20):
    flush.zero_()
    ev[2 * i].record()
    K._pack_table.repack_all()
    ev[2 * i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[2 * i].elapsed_time(ev[2 * i + 1]) * 1e3 for i in range(20))
elems = sum(e.wf.numel() + (0 if e.wd is None else e.wd.numel()) for e in K._pack_cache.values())
print(f"pack_weights_multi {name}: {len(K._pack_cache)} filters, {elems / 1e6:.1f} M packed elements, median {ts[10]:.1f} us (min {ts[0]:.1f}) "
      f"-> {elems * 2 / ts[10] / 1e3:.0f} GB/s written")
