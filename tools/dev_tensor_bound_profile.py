"""Dev tool (GPU): the generic implicit-GEMM kernel on the tensor-bound RepVGG-A0 layers (for an `ncu --set full` capture)."""
import sys

import torch

sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr

for (N, H, W, Ci, Co, k) in [(256, 14, 14, 192, 192, 3), (256, 7, 7, 1280, 1280, 3)]:
    x = torch.randn(N, H, W, Ci, device="cuda").to(torch.bfloat16)
    w = torch.randn(Co, k, k, Ci, device="cuda").to(torch.bfloat16)
    y = torch.empty(N, H, W, Co, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        lib().hb_conv2d_fprop_bf16(ptr(x), ptr(w), ptr(y), ptr(None), ptr(None), N, H, W, Ci, Co, k, k, 1, k // 2, 1, 0, 0,
                                   stream_ptr())
    torch.cuda.synchronize()
