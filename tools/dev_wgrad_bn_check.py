"""Dev harness (GPU box): wgrad, dgrad-by-fprop, and the fused BN kernels against torch on the same bf16 inputs."""
import ctypes
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from holocron_b200._lib import lib, ptr, stream_ptr

L = lib()


def check(name, got, ref, tol=2e-2):
    err = (got.float() - ref.float()).abs().max().item()
    den = ref.float().abs().max().item() + 1e-9
    nan = torch.isnan(got.float()).sum().item()
    ok = nan == 0 and err / den < tol
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.4g} ref_max={den:.4g} rel={err/den:.3g} nan={nan}", flush=True)
    return ok


def run_wgrad(N, H, W, Cin, Cout, k, stride, pad, ctas=0):
    torch.manual_seed(1)
    x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16)
    Ho = (H + 2 * pad - (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - (k - 1) - 1) // stride + 1
    dy = torch.randn(N, Cout, Ho, Wo, device="cuda").to(torch.bfloat16)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    dw = torch.full((Cout, k, k, Cin), float("nan"), device="cuda", dtype=torch.float32)
    wsb = L.hb_conv2d_wgrad_workspace_bytes(N, H, W, Cin, Cout, k, k, stride, pad, 1, ctas)
    ws = torch.empty(max(wsb // 4, 1), device='cuda')
    rc = L.hb_conv2d_wgrad_bf16(ptr(x_nhwc), ptr(dy_nhwc), ptr(dw), ptr(ws), wsb, N, H, W, Cin, Cout, k, k, stride, pad, 1, ctas,
                                stream_ptr())
    torch.cuda.synchronize()
    w = torch.zeros(Cout, Cin, k, k, device="cuda", requires_grad=True)
    y = F.conv2d(x.float(), w, stride=stride, padding=pad)
    (gw,) = torch.autograd.grad(y, w, dy.float())
    ok = check(f"wgrad rc={rc} N{N} {H}x{W} C{Cin}->{Cout} k{k} s{stride} p{pad}", dw, gw.permute(0, 2, 3, 1), 5e-3)
    return ok


def run_dgrad(N, H, W, Cin, Cout, k, stride, pad):
    """dgrad = fprop of (zero-inserted) dy with the flipped/transposed filter."""
    torch.manual_seed(2)
    w = (torch.randn(Cout, Cin, k, k, device="cuda") / (Cin * k * k) ** 0.5)
    Ho = (H + 2 * pad - (k - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - (k - 1) - 1) // stride + 1
    dy = torch.randn(N, Cout, Ho, Wo, device="cuda").to(torch.bfloat16)
    w_krsc = w.permute(0, 2, 3, 1).contiguous()  # fp32 KRSC master
    wf = torch.empty(Cout, k, k, Cin, device="cuda", dtype=torch.bfloat16)
    wd = torch.empty(Cin, k, k, Cout, device="cuda", dtype=torch.bfloat16)
    rc = L.hb_pack_conv_weights(ptr(w_krsc), ptr(wf), ptr(wd), Cout, Cin, k, k, Cin, Cin, Cout, Cout, stream_ptr())
    dy_nhwc = dy.permute(0, 2, 3, 1).contiguous()
    if stride > 1:
        dyu = torch.empty(N, H, W, Cout, device="cuda", dtype=torch.bfloat16)
        rc |= L.hb_zero_insert_bf16(ptr(dy_nhwc), ptr(dyu), N, Ho, Wo, H, W, Cout, stride, stream_ptr())
    else:
        dyu = dy_nhwc
    dx = torch.full((N, H, W, Cin), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc |= L.hb_conv2d_fprop_bf16(ptr(dyu), ptr(wd), ptr(dx), ptr(None), ptr(None), N, H, W, Cout, Cin, k, k, 1,
                                 (k - 1) - pad if stride == 1 else (k - 1) // 2 if k == 3 else 0, 1, 0, 0, stream_ptr())
    torch.cuda.synchronize()
    xz = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    y = F.conv2d(xz, wf.permute(0, 3, 1, 2).float(), stride=stride, padding=pad)
    (gx,) = torch.autograd.grad(y, xz, dy.float())
    return check(f"dgrad rc={rc} N{N} {H}x{W} C{Cin}->{Cout} k{k} s{stride} p{pad}", dx, gx.permute(0, 2, 3, 1), 2e-2)


def run_repvgg_wgrad(N, H, W, Cin, Cout):
    torch.manual_seed(6)
    x = torch.randn(N, Cin, H, W, device="cuda").to(torch.bfloat16)
    dy3 = torch.randn(N, Cout, H, W, device="cuda").to(torch.bfloat16)
    dy1 = torch.randn(N, Cout, H, W, device="cuda").to(torch.bfloat16)
    xn, d3, d1 = (t.permute(0, 2, 3, 1).contiguous() for t in (x, dy3, dy1))
    wsb = L.hb_repvgg_wgrad_workspace_bytes(N, H, W, Cin, Cout, 0)
    if wsb == 0:
        print(f"--  repvgg_wgrad N{N} {H}x{W} C{Cin}->{Cout}: not eligible")
        return True
    ws = torch.empty(wsb // 4, device="cuda")
    dw = torch.full((Cout * 10 * Cin,), float("nan"), device="cuda")
    rc = L.hb_repvgg_wgrad_bf16(ptr(xn), ptr(d3), ptr(d1), ptr(dw), ptr(ws), wsb, N, H, W, Cin, Cout, 0, stream_ptr())
    torch.cuda.synchronize()
    w3 = torch.zeros(Cout, Cin, 3, 3, device="cuda", requires_grad=True)
    w1 = torch.zeros(Cout, Cin, 1, 1, device="cuda", requires_grad=True)
    tot = (F.conv2d(x.float(), w3, padding=1) * dy3.float()).sum() + (F.conv2d(x.float(), w1) * dy1.float()).sum()
    g3, g1 = torch.autograd.grad(tot, (w3, w1))
    ok = check(f"repvgg_wgrad3 rc={rc} N{N} {H}x{W} C{Cin}->{Cout}", dw[:Cout * 9 * Cin].view(Cout, 3, 3, Cin),
               g3.permute(0, 2, 3, 1), 5e-3)
    ok &= check("   repvgg_wgrad1", dw[Cout * 9 * Cin:].view(Cout, 1, 1, Cin), g1.permute(0, 2, 3, 1), 5e-3)
    return ok


def run_dgrad_s2(N, H, W, Cin, Cout, with1x1):
    """parity-class stride-2 data gradient (+ the 1x1 stride-2 branch) vs torch autograd."""
    torch.manual_seed(4)
    w3 = torch.randn(Cout, Cin, 3, 3, device="cuda") / (Cin * 9) ** 0.5
    w1 = torch.randn(Cout, Cin, 1, 1, device="cuda") / Cin ** 0.5
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    dy3 = torch.randn(N, Ho, Wo, Cout, device="cuda").to(torch.bfloat16)
    dy1 = torch.randn(N, Ho, Wo, Cout, device="cuda").to(torch.bfloat16)
    CinD = (Cin + 15) // 16 * 16
    wcls = torch.empty(9 * CinD * Cout, device="cuda", dtype=torch.bfloat16)
    rc = L.hb_pack_dgrad_s2_weights(ptr(w3.permute(0, 2, 3, 1).contiguous()), ptr(wcls), Cout, Cin, CinD, Cout, stream_ptr())
    wd1 = torch.zeros(CinD, 1, 1, Cout, device="cuda", dtype=torch.bfloat16)
    wd1[:Cin, 0, 0, :] = w1[:, :, 0, 0].t().to(torch.bfloat16)
    dx = torch.full((N, H, W, CinD), float("nan"), device="cuda", dtype=torch.bfloat16)
    rc |= L.hb_conv2d_dgrad_s2_bf16(ptr(dy3), ptr(wcls), ptr(dy1) if with1x1 else ptr(None), ptr(wd1) if with1x1 else ptr(None),
                                    ptr(dx), N, H, W, Ho, Wo, Cout, CinD, 0, stream_ptr())
    torch.cuda.synchronize()
    xz = torch.zeros(N, Cin, H, W, device="cuda", requires_grad=True)
    y = F.conv2d(xz, w3.to(torch.bfloat16).float(), stride=2, padding=1)
    tot = (y * dy3.permute(0, 3, 1, 2).float()).sum()
    if with1x1:
        y1 = F.conv2d(xz, w1.to(torch.bfloat16).float(), stride=2, padding=0)
        tot = tot + (y1 * dy1.permute(0, 3, 1, 2).float()).sum()
    (gx,) = torch.autograd.grad(tot, xz)
    ok = check(f"dgrad_s2 rc={rc} N{N} {H}x{W} C{Cin}<-{Cout} 1x1={with1x1}", dx[..., :Cin], gx.permute(0, 2, 3, 1), 2e-2)
    if CinD != Cin:
        ok &= bool((dx[..., Cin:] == 0).all())
    return ok


def run_bn(M, C, B, act, residual):
    torch.manual_seed(3)
    dev = "cuda"
    us = [(torch.randn(M, C, device=dev) * (1 + b) + 0.5 * b).to(torch.bfloat16) for b in range(B)]
    gam = [torch.rand(C, device=dev) + 0.5 for _ in range(B)]
    bet = [torch.randn(C, device=dev) * 0.2 for _ in range(B)]
    rm = [torch.zeros(C, device=dev) for _ in range(B)]
    rv = [torch.ones(C, device=dev) for _ in range(B)]
    res = torch.randn(M, C, device=dev).to(torch.bfloat16) if residual else None
    dout = torch.randn(M, C, device=dev).to(torch.bfloat16)
    sums = torch.zeros(B, 2, C, device=dev, dtype=torch.float64)
    up = [ptr(us[b]) if b < B else ptr(None) for b in range(3)]
    rc = L.hb_bn_stats_bf16(up[0], up[1], up[2], B, M, C, ptr(sums), stream_ptr())
    mean = torch.empty(B, C, device=dev)
    rstd = torch.empty(B, C, device=dev)
    scale = torch.empty(B, C, device=dev)
    shift = torch.empty(B, C, device=dev)
    VP = ctypes.c_void_p * 3
    arr = lambda ts: VP(*[t.data_ptr() if t is not None else 0 for t in (list(ts) + [None] * 3)[:3]])
    rc |= L.hb_bn_finalize(ptr(sums), arr(gam), arr(bet), arr(rm), arr(rv), None, ptr(mean), ptr(rstd), ptr(scale), ptr(shift),
                           B, C, C, M, ctypes.c_float(1e-5), ctypes.c_float(0.1), stream_ptr())
    out = torch.full((M, C), float("nan"), device=dev, dtype=torch.bfloat16)
    rc |= L.hb_bn_act_fwd_bf16(up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(res), ptr(out), M, C, act,
                               ctypes.c_float(0.1), 0, stream_ptr())
    bsums = torch.zeros(1 + B, C, device=dev, dtype=torch.float64)
    dus = [torch.full((M, C), float("nan"), device=dev, dtype=torch.bfloat16) for _ in range(B)]
    dres = torch.full((M, C), float("nan"), device=dev, dtype=torch.bfloat16) if residual else None
    dg = torch.empty(B, C, device=dev)
    db = torch.empty(B, C, device=dev)
    dup = [ptr(dus[b]) if b < B else ptr(None) for b in range(3)]
    rc |= L.hb_bn_act_bwd_bf16(ptr(dout), up[0], up[1], up[2], B, ptr(scale), ptr(shift), ptr(mean), ptr(rstd), ptr(res),
                               ptr(bsums), dup[0], dup[1], dup[2], ptr(dres), ptr(dg), ptr(db), M, C, act,
                               ctypes.c_float(0.1), 1, 0, stream_ptr())
    torch.cuda.synchronize()
    # torch reference (fp32 math on the same bf16 inputs)
    uf = [u.float().requires_grad_(True) for u in us]
    gf = [g.clone().requires_grad_(True) for g in gam]
    bf = [b_.clone().requires_grad_(True) for b_ in bet]
    rf = res.float().requires_grad_(True) if residual else None
    z = 0
    trm, trv = [], []
    for b in range(B):
        m_, v_ = torch.zeros(C, device=dev), torch.ones(C, device=dev)
        z = z + F.batch_norm(uf[b], m_, v_, gf[b], bf[b], True, 0.1, 1e-5)
        trm.append(m_)
        trv.append(v_)
    if residual:
        z = z + rf
    acts = {0: lambda t: t, 1: torch.relu, 2: F.relu6, 3: F.silu, 4: lambda t: F.leaky_relu(t, 0.1), 5: F.mish,
            6: lambda t: 0.5 * t * (t + 2).clamp(0, 2)}
    o = acts[act](z)
    o.backward(dout.float())
    ok = rc == 0
    ok &= check(f"bn fwd M{M} C{C} B{B} act{act} res{residual} rc={rc}", out, o, 1e-2)
    for b in range(B):
        ok &= check(f"   running_mean[{b}]", rm[b], trm[b], 1e-4)
        ok &= check(f"   running_var[{b}]", rv[b], trv[b], 1e-4)
        ok &= check(f"   du[{b}]", dus[b], uf[b].grad, 1.5e-2)
        ok &= check(f"   dgamma[{b}]", dg[b], gf[b].grad, 2e-3)
        ok &= check(f"   dbeta[{b}]", db[b], bf[b].grad, 2e-3)
    if residual:
        ok &= check("   dres", dres, rf.grad, 1e-2)
    return ok


def main():
    print(torch.cuda.get_device_name(0))
    ok = True
    for cfg in [(2, 16, 16, 64, 64, 1, 1, 0), (2, 16, 16, 64, 64, 3, 1, 1), (2, 14, 14, 48, 48, 3, 1, 1),
                (4, 28, 28, 96, 96, 3, 1, 1), (3, 14, 14, 192, 192, 3, 1, 1), (2, 28, 28, 48, 96, 3, 2, 1),
                (2, 28, 28, 48, 96, 1, 2, 0), (2, 7, 7, 192, 1280, 3, 1, 1), (8, 7, 7, 1280, 1280, 3, 1, 1),
                (2, 56, 56, 8, 48, 3, 2, 1), (16, 56, 56, 48, 48, 3, 1, 1), (16, 56, 56, 48, 48, 1, 1, 0),
                (3, 112, 112, 48, 48, 3, 1, 1), (2, 30, 20, 32, 48, 3, 1, 1), (2, 9, 11, 16, 16, 3, 1, 1),
                (5, 28, 28, 64, 96, 3, 1, 1), (2, 17, 33, 8, 32, 3, 1, 1), (300, 14, 14, 48, 48, 3, 1, 1),
                (2, 126, 126, 24, 40, 3, 1, 1), (1, 8, 8, 64, 64, 3, 1, 1), (2, 16, 16, 72, 200, 3, 1, 1),
                (2, 20, 12, 256, 256, 3, 1, 1), (40, 14, 14, 192, 192, 3, 1, 1)]:
        try:
            ok &= run_wgrad(*cfg)
        except Exception as e:  # noqa: BLE001
            print("EXC wgrad", cfg, repr(e)); ok = False
    for cfg in [(2, 16, 16, 64, 64, 3, 1, 1), (2, 14, 14, 48, 48, 3, 1, 1), (2, 14, 14, 48, 48, 1, 1, 0),
                (2, 28, 28, 48, 96, 3, 2, 1), (2, 28, 28, 48, 96, 1, 2, 0), (2, 14, 14, 192, 1280, 3, 2, 1)]:
        try:
            ok &= run_dgrad(*cfg)
        except Exception as e:  # noqa: BLE001
            print("EXC dgrad", cfg, repr(e)); ok = False
    for cfg in [(3, 112, 112, 48, 48), (5, 28, 28, 96, 96), (40, 14, 14, 192, 192), (2, 30, 20, 32, 48), (2, 16, 16, 72, 200),
                (2, 9, 11, 16, 16), (300, 14, 14, 48, 48)]:
        try:
            ok &= run_repvgg_wgrad(*cfg)
        except Exception as e:  # noqa: BLE001
            print("EXC repvgg_wgrad", cfg, repr(e)); ok = False
    for cfg in [(2, 16, 16, 48, 48, True), (2, 16, 16, 48, 96, False), (3, 15, 9, 24, 32, True), (2, 14, 14, 192, 1280, True),
                (2, 7, 7, 64, 64, False), (4, 56, 56, 48, 96, True), (1, 2, 2, 16, 16, True), (2, 224, 224, 8, 48, False)]:
        try:
            ok &= run_dgrad_s2(*cfg)
        except Exception as e:  # noqa: BLE001
            print("EXC dgrad_s2", cfg, repr(e)); ok = False
    for cfg in [(1000, 48, 3, 1, False), (4096, 64, 1, 1, False), (777, 1280, 2, 1, False), (2048, 96, 3, 0, True),
                (2048, 320, 1, 3, False), (2048, 32, 1, 4, True), (2048, 64, 1, 5, False), (512, 8, 2, 6, False),
                (3000, 192, 1, 2, True)]:
        try:
            ok &= run_bn(*cfg)
        except Exception as e:  # noqa: BLE001
            print("EXC bn", cfg, repr(e)); ok = False
    print("ALL OK" if ok else "SOME BAD")
    # timing: wgrad of the big layers
    for shape in [(256, 112, 112, 48, 48), (256, 56, 56, 48, 48), (256, 28, 28, 96, 96), (256, 14, 14, 192, 192),
                  (256, 7, 7, 1280, 1280)]:
        N, H, W, Ci, Co = shape
        x = torch.randn(N, H, W, Ci, device="cuda").to(torch.bfloat16)
        dy = torch.randn(N, H, W, Co, device="cuda").to(torch.bfloat16)
        dw = torch.empty(Co, 3, 3, Ci, device="cuda")
        wsb = L.hb_conv2d_wgrad_workspace_bytes(N, H, W, Ci, Co, 3, 3, 1, 1, 1, 0)
        ws = torch.empty(max(wsb // 4, 1), device='cuda')
        args = (ptr(x), ptr(dy), ptr(dw), ptr(ws), wsb, N, H, W, Ci, Co, 3, 3, 1, 1, 1, 0, stream_ptr())
        for _ in range(3):
            L.hb_conv2d_wgrad_bf16(*args)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            L.hb_conv2d_wgrad_bf16(*args)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        flops = 2 * N * H * W * Co * Ci * 9
        print(f"wgrad time {shape}: {ms:.3f} ms {flops/ms/1e9:.1f} TFLOP/s {(x.numel()+dy.numel())*2/ms/1e6:.0f} GB/s")
    # timing: bn kernels on the biggest activation
    M, C = 256 * 112 * 112, 48
    u = [torch.randn(M, C, device="cuda").to(torch.bfloat16) for _ in range(3)]
    out = torch.empty_like(u[0])
    sums = torch.zeros(3, 2, C, device="cuda", dtype=torch.float64)
    scale = torch.ones(3, C, device="cuda"); shift = torch.zeros(3, C, device="cuda")
    for name, fn in [("stats3", lambda: L.hb_bn_stats_bf16(ptr(u[0]), ptr(u[1]), ptr(u[2]), 3, M, C, ptr(sums), stream_ptr())),
                     ("fwd3", lambda: L.hb_bn_act_fwd_bf16(ptr(u[0]), ptr(u[1]), ptr(u[2]), 3, ptr(scale), ptr(shift), ptr(None), ptr(out), M, C, 1, ctypes.c_float(0.0), 0, stream_ptr()))]:
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nb = (3 if name == "stats3" else 4) * M * C * 2
        print(f"bn {name}: {ms:.3f} ms  {nb/ms/1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
