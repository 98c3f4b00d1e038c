"""Dev tool (GPU): device time / effective HBM bandwidth of the fused BN kernels on the RepVGG-A0 batch-256 shapes."""
import ctypes
import sys

import torch

sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr

L = lib()
dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


for (M, C, B, residual) in [(256 * 112 * 112, 48, 2, False), (256 * 56 * 56, 48, 3, False), (256 * 28 * 28, 96, 3, False),
                            (256 * 14 * 14, 192, 3, False), (256 * 7 * 7, 1280, 2, False), (256 * 28 * 28, 128, 1, True)]:
    us = [torch.randn(M, C, device=dev).to(torch.bfloat16) for _ in range(B)]
    res = torch.randn(M, C, device=dev).to(torch.bfloat16) if residual else None
    dout = torch.randn(M, C, device=dev).to(torch.bfloat16)
    up = [ptr(us[b]) if b < B else ptr(None) for b in range(3)]
    sums = torch.zeros(B, 2, C, device=dev, dtype=torch.float64)
    mean, rstd, scale, shift = (torch.rand(B, C, device=dev) + 0.5 for _ in range(4))
    out = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    bsums = torch.zeros(1 + B, C, device=dev, dtype=torch.float64)
    dus = [torch.empty(M, C, device=dev, dtype=torch.bfloat16) for _ in range(B)]
    dres = torch.empty(M, C, device=dev, dtype=torch.bfloat16) if residual else None
    dg, db = torch.empty(B, C, device=dev), torch.empty(B, C, device=dev)
    dup = [ptr(dus[b]) if b < B else ptr(None) for b in range(3)]
    tb = M * C * 2 / 1e9   # GB per tensor
    t = timeit(lambda: L.hb_bn_stats_bf16(up[0], up[1], up[2], B, M, C, ptr(sums), stream_ptr()))
    line = f"M={M} C={C} B={B} res={int(residual)}: stats {t:.3f} ms {B * tb / t * 1e3:.0f} GB/s"
    t = timeit(lambda: L.hb_bn_act_fwd_bf16(up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(res), ptr(out), M, C, 1,
                                            ctypes.c_float(0.1), 0, stream_ptr()))
    line += f" | fwd {t:.3f} ms {(B + 1 + int(residual)) * tb / t * 1e3:.0f} GB/s"
    t = timeit(lambda: L.hb_bn_act_bwd_bf16(ptr(dout), up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                            ptr(res), ptr(bsums), dup[0], dup[1], dup[2], ptr(dres), ptr(dg), ptr(db), M, C, 1,
                                            ctypes.c_float(0.1), 1, 0, stream_ptr()))
    nbytes = (2 * (B + 1 + int(residual)) + B + int(residual)) * tb
    line += f" | bwd {t:.3f} ms {nbytes / t * 1e3:.0f} GB/s"
    t2 = timeit(lambda: L.hb_bn_act_bwd_bf16(ptr(dout), up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(mean), ptr(rstd),
                                             ptr(res), ptr(bsums), dup[0], dup[1], dup[2], ptr(dres), ptr(None), ptr(None), M, C,
                                             1, ctypes.c_float(0.1), 0, 0, stream_ptr()))
    nb_apply = ((B + 1 + int(residual)) + B + int(residual)) * tb
    line += f" (apply {t2:.3f} ms {nb_apply / t2 * 1e3:.0f} GB/s, reduce {t - t2:.3f} ms {(B + 1 + int(residual)) * tb / max(t - t2, 1e-6) * 1e3:.0f} GB/s)"
    print(line, flush=True)
    del us, res, dout, out, dus, dres
