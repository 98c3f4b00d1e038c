"""Oracle restatements of the optimizer updates (reference: holocron/optim/{adabelief,lamb,tadam}.py).

Each function performs ONE step in place on plain tensors (fp32, CPU) and returns nothing, mirroring the
reference's per-tensor update formulas, quirks included.
"""
import math
from typing import Optional, Tuple

import torch
from torch import Tensor


@torch.no_grad()
def adabelief_step(p: Tensor, g: Tensor, m: Tensor, s: Tensor, step: int, lr: float, beta1: float, beta2: float,
                   eps: float, weight_decay: float = 0.0, amsgrad: bool = False,
                   s_max: Optional[Tensor] = None) -> None:
    """reference optim/adabelief.py:121-167. No +eps inside the belief EMA (unlike the paper)."""
    if weight_decay != 0:
        g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)          # in-place: the CPU baseline times this function
    r = g - m
    s.mul_(beta2).addcmul_(r, r, value=1 - beta2)
    second = s
    if amsgrad:
        torch.maximum(s_max, s, out=s_max)
        second = s_max
    denom = second.sqrt().div_(math.sqrt(1 - beta2**step)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / (1 - beta1**step)))


@torch.no_grad()
def lamb_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, lr: float, beta1: float, beta2: float, eps: float,
              weight_decay: float = 0.0, scale_clip: Tuple[float, float] = (0.0, 10.0)) -> float:
    """reference optim/lamb.py:79-137: Adam moments without bias correction, LARS-style trust ratio.
    Returns the local learning-rate multiplier."""
    m.copy_(beta1 * m + (1 - beta1) * g)
    v.copy_(beta2 * v + (1 - beta2) * g * g)
    update = m / (v.sqrt() + eps)
    if weight_decay != 0:
        update = update + weight_decay * p
    p_norm = p.pow(2).sum().sqrt()
    u_norm = update.pow(2).sum().sqrt()
    phi = p_norm.clamp(*scale_clip)
    local_lr = 1.0 if (phi == 0 or u_norm == 0) else float(phi / u_norm)
    p.sub_(lr * local_lr * update)
    return local_lr


@torch.no_grad()
def tadam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, W: Tensor, step: int, lr: float, beta1: float,
               beta2: float, eps: float, weight_decay: float = 0.0, dof: Optional[float] = None,
               amsgrad: bool = False, v_max: Optional[Tensor] = None) -> None:
    """reference optim/tadam.py:160-212. ``W`` is the 1-element running weight sum (init beta1/(1-beta1))."""
    n = p.numel()
    d = float(n) if dof is None else float(dof)
    if weight_decay != 0:
        g = g + weight_decay * p
    w = ((g - m) ** 2 / (v + eps)).sum()
    w = (d + n) / (w + d)
    m.copy_(m * (W / (W + w)) + (w * g) / (W + w))
    W.copy_(W * ((2 * beta1 - 1) / beta1) + w)
    v.copy_(beta2 * v + (1 - beta2) * g * g)
    second = v
    if amsgrad:
        v_max.copy_(torch.maximum(v_max, v))
        second = v_max
    denom = second.sqrt() / math.sqrt(1 - beta2**step) + eps
    p.sub_((lr / (1 - beta1**step)) * m / denom)


@torch.no_grad()
def adamp_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float, eps: float,
               weight_decay: float = 0.0, delta: float = 0.1, amsgrad: bool = False, v_max: Optional[Tensor] = None) -> None:
    """reference optim/adamp.py:144-191: Adam moments (L2 decay folded into the gradient), then - when the gradient is almost
    orthogonal to the weights, cosine_similarity(p, g) < delta / sqrt(numel) - the update's component along the weights is
    projected out before it is applied."""
    import torch.nn.functional as F
    bc1, bc2 = 1 - beta1**step, 1 - beta2**step
    if weight_decay != 0:
        g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    second = v
    if amsgrad:
        torch.maximum(v_max, v, out=v_max)
        second = v_max
    denom = (second.sqrt() / math.sqrt(bc2)).add_(eps)
    pt = m / bc1 / denom
    if F.cosine_similarity(p.view(1, -1), g.view(1, -1)).max() < delta / math.sqrt(p.numel()):
        pn = p / p.norm().add_(eps)
        pt -= (pn * pt).sum() * pn
    p.add_(pt, alpha=-lr)
