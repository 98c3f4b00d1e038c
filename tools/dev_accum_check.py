"""Dev: fused 3x3 + 1x1 + identity accumulate kernel vs torch."""
import sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr
torch.manual_seed(0)
ok = True
for (N, H, W, C, Co, nextra) in [(2, 14, 14, 48, 48, 2), (2, 56, 56, 48, 48, 2), (3, 112, 112, 48, 48, 2), (2, 28, 28, 64, 64, 1),
                                 (2, 30, 20, 32, 48, 1), (2, 112, 112, 64, 64, 2)]:
    x = torch.randn(N, H, W, C, device="cuda").bfloat16()
    w = (torch.randn(Co, 3, 3, C, device="cuda") * 0.05).bfloat16()
    x1 = torch.randn(N, H, W, C, device="cuda").bfloat16()
    w1 = (torch.randn(Co, 1, 1, C, device="cuda") * 0.1).bfloat16()
    x2 = torch.randn(N, H, W, C, device="cuda").bfloat16()
    w2 = torch.eye(Co, C, device="cuda").bfloat16().reshape(Co, 1, 1, C).contiguous()
    y = torch.full((N, H, W, Co), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc = lib().hb_conv3x3_accum_bf16(ptr(x), ptr(w), ptr(x1), ptr(w1), ptr(x2), ptr(w2), nextra, ptr(y), N, H, W, C, Co, 0, stream_ptr())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), padding=1)
    ref = ref + F.conv2d(x1.float().permute(0, 3, 1, 2), w1.float().permute(0, 3, 1, 2))
    if nextra == 2:
        ref = ref + F.conv2d(x2.float().permute(0, 3, 1, 2), w2.float().permute(0, 3, 1, 2))
    ref = ref.permute(0, 2, 3, 1)
    err = (y.float() - ref).abs().max().item() / ref.abs().max().item()
    nan = torch.isnan(y.float()).sum().item()
    good = rc == 0 and nan == 0 and err < 1e-2
    ok &= good
    print(("OK " if good else "BAD"), (N, H, W, C, Co, nextra), "rc", rc, "nan", nan, "rel", round(err, 5), flush=True)
print("ALL OK" if ok else "SOME BAD")
N, H, W, C = 256, 112, 112, 48
x = [torch.randn(N, H, W, C, device="cuda").bfloat16() for _ in range(3)]
w = (torch.randn(C, 3, 3, C, device="cuda") * 0.05).bfloat16()
w1 = (torch.randn(C, 1, 1, C, device="cuda") * 0.1).bfloat16()
w2 = torch.eye(C, C, device="cuda").bfloat16().reshape(C, 1, 1, C).contiguous()
y = torch.empty(N, H, W, C, device="cuda", dtype=torch.bfloat16)
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for ne in (0, 1, 2):
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib().hb_conv3x3_accum_bf16(ptr(x[0]), ptr(w), ptr(x[1]), ptr(w1), ptr(x[2]), ptr(w2), ne, ptr(y), N, H, W, C, C, 0, stream_ptr())
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    byts = N * H * W * C * 2 * (2 + ne)
    print(f"accum nextra={ne} 112^2 C48: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s")
