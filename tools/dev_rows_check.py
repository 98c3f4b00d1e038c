"""Dev harness: row-window conv kernel + new epilogue vs torch; timing of the RepVGG stage-0/1 shapes."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr
sys.path.insert(0, "tools")
from dev_conv_check import run_fprop

ok = True
for cfg in [(2, 16, 16, 64, 64, 3, 1, 1), (2, 14, 14, 48, 48, 3, 1, 1), (3, 28, 28, 48, 48, 3, 1, 1), (2, 56, 56, 48, 48, 3, 1, 1),
            (2, 112, 112, 48, 48, 3, 1, 1), (1, 30, 30, 64, 64, 3, 1, 1), (2, 17, 23, 16, 32, 3, 1, 1), (2, 9, 11, 8, 16, 3, 1, 1),
            (5, 112, 112, 64, 64, 3, 1, 1), (2, 56, 56, 96, 48, 3, 1, 1), (2, 126, 126, 48, 48, 3, 1, 1)]:
    ok &= run_fprop(*cfg)
ok &= run_fprop(2, 28, 28, 48, 48, 3, 1, 1, bias=True, act=1)
ok &= run_fprop(2, 28, 28, 64, 64, 3, 1, 1, residual=True, act=1)
ok &= run_fprop(2, 16, 16, 64, 128, 1, 1, 0)
ok &= run_fprop(2, 7, 7, 192, 1280, 3, 1, 1)
ok &= run_fprop(2, 28, 28, 48, 96, 3, 2, 1)
print("ALL OK" if ok else "SOME BAD")
for shape in [(256, 112, 112, 48, 48, 3), (256, 56, 56, 48, 48, 3), (256, 112, 112, 64, 64, 3), (256, 112, 112, 48, 48, 1), (256, 28, 28, 96, 96, 3)]:
    N, H, W, Ci, Co, k = shape
    x = torch.randn(N, H, W, Ci, device="cuda").to(torch.bfloat16)
    w = torch.randn(Co, k, k, Ci, device="cuda").to(torch.bfloat16)
    y = torch.empty(N, H, W, Co, device="cuda", dtype=torch.bfloat16)
    flush = torch.empty(256 * 1024 * 1024, device="cuda", dtype=torch.uint8)
    args = (ptr(x), ptr(w), ptr(y), ptr(None), ptr(None), N, H, W, Ci, Co, k, k, 1, k // 2, 1, 0, 0, stream_ptr())
    for _ in range(3):
        lib().hb_conv2d_fprop_bf16(*args)
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); lib().hb_conv2d_fprop_bf16(*args); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    byts = (N * H * W * (Ci + Co) + Co * Ci * k * k) * 2
    print(f"time {shape}: {ms:.3f} ms  {2*N*H*W*Co*Ci*k*k/ms/1e9:.0f} TFLOP/s  {byts/ms/1e6:.0f} GB/s (L2 flushed)")
