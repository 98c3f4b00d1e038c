"""focal_loss / poly_loss / dice_loss on the fused CUDA kernels (holocron_b200/csrc/losses.cu)."""
import ctypes
from typing import Optional

import torch
from torch import Tensor

from .._lib import check, dtype_code, lib, ptr, require_cuda, stream_ptr

_cf = ctypes.c_float
_RED = {"none": 0, "mean": 1, "sum": 2}


def _nks(x: Tensor):
    n, k = x.shape[0], x.shape[1]
    s = 1
    for d in x.shape[2:]:
        s *= d
    return n, k, s


def _weight(weight, x: Tensor) -> Optional[Tensor]:
    if not isinstance(weight, Tensor):
        return None
    return weight.detach().to(device=x.device, dtype=torch.float32).contiguous()


class _HardLossFn(torch.autograd.Function):
    """kind 0: focal, 1: poly-1; hard (int64) targets."""

    @staticmethod
    def forward(ctx, x: Tensor, target: Tensor, weight: Optional[Tensor], ignore_index: int, reduction: int, kind: int,
                gamma: float, eps: float) -> Tensor:
        require_cuda(x, target)
        xc = x.contiguous()
        tc = target.contiguous().view(-1)
        if tc.dtype != torch.long:
            tc = tc.long()
        n, k, s = _nks(xc)
        if tc.numel() != n * s:
            raise ValueError("target shape does not match the input's (N, ...) dims")
        L = lib()
        loss_pos = torch.empty(n * s, device=x.device, dtype=torch.float32)
        partials = torch.empty(2 * L.hb_loss_max_partials(), device=x.device, dtype=torch.float64)
        fwd_out = torch.empty(3, device=x.device, dtype=torch.float32)
        check(L.hb_cls_loss_hard_fwd(ptr(xc), ptr(tc), ptr(weight), ptr(loss_pos), ptr(partials), ptr(fwd_out), n, k, s,
                                     int(ignore_index), kind, _cf(gamma), _cf(eps), dtype_code(xc), stream_ptr()),
              "hb_cls_loss_hard_fwd")
        ctx.save_for_backward(xc, tc, weight, fwd_out)
        ctx.cfg = (n, k, s, int(ignore_index), reduction, kind, gamma, eps)
        if reduction == 1:
            return fwd_out[2].to(x.dtype)
        if reduction == 2:
            return fwd_out[0].to(x.dtype)
        return loss_pos.to(x.dtype)

    @staticmethod
    def backward(ctx, gout: Tensor):
        xc, tc, weight, fwd_out = ctx.saved_tensors
        n, k, s, ignore_index, reduction, kind, gamma, eps = ctx.cfg
        g = gout.detach().float().contiguous().view(-1)
        dx = torch.empty_like(xc)
        check(lib().hb_cls_loss_hard_bwd(ptr(xc), ptr(tc), ptr(weight), ptr(g), ptr(fwd_out), ptr(dx), n, k, s,
                                         ignore_index, kind, _cf(gamma), _cf(eps), reduction, dtype_code(xc),
                                         stream_ptr()), "hb_cls_loss_hard_bwd")
        return dx, None, None, None, None, None, None, None


class _PolySoftFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, target: Tensor, weight: Optional[Tensor], ignore_index: int, reduction: int,
                eps: float) -> Tensor:
        require_cuda(x, target)
        xc = x.contiguous()
        tc = target.to(xc.dtype).contiguous()
        n, k, s = _nks(xc)
        L = lib()
        loss_pos = torch.empty(n * s, device=x.device, dtype=torch.float32)
        partials = torch.empty(2 * L.hb_loss_max_partials(), device=x.device, dtype=torch.float64)
        fwd_out = torch.empty(3, device=x.device, dtype=torch.float32)
        check(L.hb_poly_soft_fwd(ptr(xc), ptr(tc), ptr(weight), ptr(loss_pos), ptr(partials), ptr(fwd_out), n, k, s,
                                 int(ignore_index), _cf(eps), dtype_code(xc), stream_ptr()), "hb_poly_soft_fwd")
        ctx.save_for_backward(xc, tc, weight)
        ctx.cfg = (n, k, s, int(ignore_index), reduction, eps)
        if reduction == 1:
            return fwd_out[2].to(x.dtype)
        if reduction == 2:
            return fwd_out[0].to(x.dtype)
        return loss_pos.to(x.dtype).view(x.shape[0], *x.shape[2:])

    @staticmethod
    def backward(ctx, gout: Tensor):
        xc, tc, weight = ctx.saved_tensors
        n, k, s, ignore_index, reduction, eps = ctx.cfg
        g = gout.detach().float().contiguous().view(-1)
        dx = torch.empty_like(xc)
        check(lib().hb_poly_soft_bwd(ptr(xc), ptr(tc), ptr(weight), ptr(g), ptr(dx), n, k, s, ignore_index, _cf(eps),
                                     reduction, dtype_code(xc), stream_ptr()), "hb_poly_soft_bwd")
        return dx, None, None, None, None, None


def _check_reduction(reduction: str) -> int:
    # the reference treats every value other than "sum" / "mean" as "none"
    return _RED.get(reduction, 0)


def focal_loss(x: Tensor, target: Tensor, weight: Optional[Tensor] = None, ignore_index: int = -100,
               reduction: str = "mean", gamma: float = 2.0) -> Tensor:
    """Focal loss — mirrors holocron/nn/functional.py:59-113, quirks included: the class weight scales
    ``log p_t`` only, ``ignore_index`` is honoured only inside ``[0, K)``, ``'mean'`` averages over the non-ignored
    positions and ``'none'`` is shaped like ``target``. Out-of-range targets give NaN (the reference raises from
    ``gather``; detecting it here would cost a device synchronisation)."""
    out = _HardLossFn.apply(x, target, _weight(weight, x), ignore_index, _check_reduction(reduction), 0, float(gamma), 0.0)
    if _check_reduction(reduction) == 0:
        return out.view(*target.shape)
    return out


def poly_loss(x: Tensor, target: Tensor, eps: float = 2.0, weight: Optional[Tensor] = None, ignore_index: int = -100,
              reduction: str = "mean") -> Tensor:
    """Poly-1 loss — mirrors holocron/nn/functional.py:540-613 for hard (``int64``, ``(N, ...)``) and soft
    (``(N, K, ...)``) targets. ``reduction='none'`` returns the flat ``(N*...,)`` vector for hard targets and
    ``(N, ...)`` for soft ones, as the reference does."""
    if target.ndim == x.ndim - 1:
        if target.dtype != torch.long:
            raise TypeError("target dtype is expected to be torch.int64")
        return _HardLossFn.apply(x, target, _weight(weight, x), ignore_index, _check_reduction(reduction), 1, 0.0,
                                 float(eps))
    if target.ndim != x.ndim or target.shape[0] != x.shape[0] or target.shape[1] != x.shape[1]:
        raise ValueError("invalid target shape")
    if isinstance(weight, Tensor) and x.ndim != 2:
        # the reference broadcasts weight.reshape(1, -1) against the LAST dim there, which is not a class weighting
        raise NotImplementedError("class weights with soft targets are only defined for (N, K) inputs")
    return _PolySoftFn.apply(x, target, _weight(weight, x), ignore_index, _check_reduction(reduction), float(eps))


class _DiceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor, target: Tensor, weight: Optional[Tensor], gamma: float, eps: float) -> Tensor:
        require_cuda(x, target)
        if x.ndim < 3:
            raise IndexError("Dimension out of range (dice_loss expects (N, K, ...) inputs with >= 3 dims)")
        xc = x.contiguous()
        tc = target.to(xc.dtype).contiguous()
        n, k, s = _nks(xc)
        sums = torch.empty(lib().hb_dice_scratch_doubles(k), device=x.device, dtype=torch.float64)
        out = torch.empty(1, device=x.device, dtype=torch.float32)
        coef = torch.empty(2 * k, device=x.device, dtype=torch.float32)
        check(lib().hb_dice_fwd(ptr(xc), ptr(tc), ptr(weight), ptr(sums), ptr(out), ptr(coef), n, k, s, _cf(gamma),
                                _cf(eps), dtype_code(xc), stream_ptr()), "hb_dice_fwd")
        ctx.save_for_backward(tc, coef)
        ctx.cfg = (n, k, s)
        return out[0].to(x.dtype)

    @staticmethod
    def backward(ctx, gout: Tensor):
        tc, coef = ctx.saved_tensors
        n, k, s = ctx.cfg
        g = gout.detach().float().contiguous().view(-1)
        dx = torch.empty_like(tc)
        check(lib().hb_dice_bwd(ptr(tc), ptr(coef), ptr(g), ptr(dx), n, k, s, dtype_code(tc), stream_ptr()), "hb_dice_bwd")
        return dx, None, None, None, None


def dice_loss(x: Tensor, target: Tensor, weight: Optional[Tensor] = None, gamma: float = 1.0, eps: float = 1e-8) -> Tensor:
    """Dice loss on probabilities — mirrors holocron/nn/functional.py:503-537:
    ``1 - (1 + 1/gamma) * mean_k[(gamma*sum(x*t) + eps) / (sum(x + gamma*t) + eps)]`` with the sums taken jointly over
    batch and space. Two streaming reductions per class in one pass instead of 4 full-tensor temporaries."""
    return _DiceFn.apply(x, target, _weight(weight, x), float(gamma), float(eps))
