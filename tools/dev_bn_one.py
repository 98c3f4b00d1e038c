"""Dev tool (GPU): one launch of each fused BN kernel on the largest RepVGG-A0 shape (for ncu captures)."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr

L = lib()
dev = "cuda"
M, C, B = 256 * 112 * 112, 48, 2
us = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(B)]
dout = torch.randn(M, C, device=dev).to(torch.bfloat16)
up = [ptr(us[b]) if b < B else ptr(None) for b in range(3)]
sums = torch.zeros(B, 2, C, device=dev, dtype=torch.float64)
mean, rstd, scale, shift = (torch.rand(B, C, device=dev) + 0.5 for _ in range(4))
out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
bsums = torch.zeros(1 + B, C, device=dev, dtype=torch.float64)
dus = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(B)]
dg, db = torch.empty(B, C, device=dev), torch.empty(B, C, device=dev)
dup = [ptr(dus[b]) if b < B else ptr(None) for b in range(3)]
for _ in range(2):
    L.hb_bn_stats_bf16(up[0], up[1], up[2], B, M, C, ptr(sums), stream_ptr())
    L.hb_bn_act_fwd_bf16(up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(None), ptr(out), M, C, 1, ctypes.c_float(0.1), 0,
                         stream_ptr())
    L.hb_bn_act_bwd_bf16(ptr(dout), up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(None), ptr(bsums),
                         dup[0], dup[1], dup[2], ptr(None), ptr(dg), ptr(db), M, C, 1, ctypes.c_float(0.1), 1, 0, stream_ptr())
    out2 = out + out   # torch elementwise reference point (2 reads + 1 write of the same size)
torch.cuda.synchronize()
