"""norm_conv2d / add2d on the patch cross-correlation kernels (holocron_b200/csrc/xcorr.cu)."""
import ctypes
from typing import Optional, Tuple, Union

import torch
from torch import Tensor

from .._lib import check, lib, ptr, require_cuda, stream_ptr

_cf = ctypes.c_float


def _single(v: Union[int, Tuple[int, int]], what: str) -> int:
    if isinstance(v, (tuple, list)):
        if len(v) != 2 or v[0] != v[1]:
            raise NotImplementedError(f"only symmetric {what} is supported by the fused kernel")
        return int(v[0])
    return int(v)


class _XcorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, pad: int, dil: int, mode: int,
                normalize: bool, eps: float) -> Tensor:
        require_cuda(x, weight)
        if x.ndim != 4 or weight.ndim != 4:
            raise ValueError("expected (N, C, H, W) input and (Cout, Cin, kh, kw) weight")
        n, cin, h, w = x.shape
        cout, cin_w, kh, kw = weight.shape
        if cin_w != cin:
            # the reference ignores `groups`: a grouped weight makes its matmul fail with a shape error
            raise RuntimeError(f"weight expects {cin_w} input channels but the input has {cin} (groups are ignored)")
        x32 = x.detach().float().contiguous()
        w32 = weight.detach().float().contiguous()
        b32 = None if bias is None else bias.detach().float().contiguous()
        ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
        wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        out = torch.empty((n, cout, ho, wo), device=x.device, dtype=torch.float32)
        mean = torch.empty(n * ho * wo if normalize else 1, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        check(lib().hb_xcorr2d_fwd(ptr(x32), ptr(w32), ptr(b32), ptr(out), ptr(mean), ptr(rstd), n, cin, h, w, cout, kh, kw,
                                   stride, pad, dil, mode, int(normalize), _cf(eps), stream_ptr()), "hb_xcorr2d_fwd")
        ctx.save_for_backward(x32, w32, mean, rstd)
        ctx.cfg = (stride, pad, dil, mode, normalize, eps, bias is not None, x.dtype, weight.dtype)
        return out.to(x.dtype)

    @staticmethod
    def backward(ctx, gout: Tensor):
        x32, w32, mean, rstd = ctx.saved_tensors
        stride, pad, dil, mode, normalize, eps, has_bias, xdt, wdt = ctx.cfg
        n, cin, h, w = x32.shape
        cout, _, kh, kw = w32.shape
        g = gout.detach().float().contiguous()
        L = lib()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            if normalize or mode == 0:
                # same situation as the reference, whose in-place slice normalisation makes autograd raise
                raise RuntimeError("the input gradient of a slice-normalised cross-correlation is not defined by the "
                                   "reference (its in-place normalisation breaks autograd); only add2d without "
                                   "normalize_slices propagates to the input")
            dx = torch.empty_like(x32)
            check(L.hb_add2d_dgrad(ptr(x32), ptr(w32), ptr(g), ptr(dx), n, cin, h, w, cout, kh, kw, stride, pad, dil,
                                   stream_ptr()), "hb_add2d_dgrad")
            dx = dx.to(xdt)
        if ctx.needs_input_grad[1]:
            dw = torch.empty_like(w32)
            check(L.hb_xcorr2d_wgrad(ptr(x32), ptr(w32), ptr(g), ptr(mean), ptr(rstd), ptr(dw), n, cin, h, w, cout, kh, kw,
                                     stride, pad, dil, mode, int(normalize), _cf(eps), stream_ptr()), "hb_xcorr2d_wgrad")
            dw = dw.to(wdt)
        if has_bias and ctx.needs_input_grad[2]:
            db = g.sum((0, 2, 3))
        return dx, dw, db, None, None, None, None, None, None


def norm_conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: Union[int, Tuple[int, int]] = 1,
                padding: Union[int, Tuple[int, int]] = 0, dilation: Union[int, Tuple[int, int]] = 1, groups: int = 1,
                eps: float = 1e-14) -> Tensor:
    """Normalised convolution — mirrors holocron/nn/functional.py:378-413: every im2col patch (the whole
    ``Cin*kh*kw`` vector, zero padding included) is standardised with its biased variance, then correlated with the
    filters. ``groups`` is accepted and ignored, as in the reference. No im2col tensor is materialised."""
    return _XcorrFn.apply(x, weight, bias, _single(stride, "stride"), _single(padding, "padding"),
                          _single(dilation, "dilation"), 0, True, float(eps))


def add2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: Union[int, Tuple[int, int]] = 1,
          padding: Union[int, Tuple[int, int]] = 0, dilation: Union[int, Tuple[int, int]] = 1, groups: int = 1,
          normalize_slices: bool = False, eps: float = 1e-14) -> Tensor:
    """AdderNet layer ``-sum_k |patch_k - w_k|`` — mirrors holocron/nn/functional.py:426-462. The reference broadcasts
    an ``N x L x Cout x K`` tensor; here the L1 distances are accumulated tile by tile in shared memory."""
    return _XcorrFn.apply(x, weight, bias, _single(stride, "stride"), _single(padding, "padding"),
                          _single(dilation, "dilation"), 1, bool(normalize_slices), float(eps))
