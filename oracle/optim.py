"""Oracle restatements of the optimizer updates (reference: holocron/optim/{adabelief,lamb,tadam,adamp,adan,ademamix,lars,ralars,wrapper}.py).

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


@torch.no_grad()
def adan_step(p: Tensor, g: Tensor, prev_g: Tensor, m: Tensor, v: Tensor, n: Tensor, step: int, lr: float, beta1: float,
              beta2: float, beta3: float, eps: float, weight_decay: float = 0.0, amsgrad: bool = False,
              n_max: Optional[Tensor] = None) -> None:
    """reference optim/adan.py:145-199. ``prev_g`` is read and never written (the reference never updates
    state['prev_grad']); the update mixes ``beta2 * v / bc2``; with weight decay ``p /= 1 + wd * lr`` after the step."""
    bc1, bc2, bc3 = 1 - beta1**step, 1 - beta2**step, 1 - beta3**step
    if weight_decay != 0:
        g = g + weight_decay * p
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    dg = g - prev_g
    v.mul_(beta2).add_(dg, alpha=1 - beta2)
    tmp = g + beta2 * dg
    n.mul_(beta3).addcmul_(tmp, tmp, value=1 - beta3)
    second = n
    if amsgrad:
        torch.maximum(n_max, n, out=n_max)
        second = n_max
    denom = (second.sqrt() / math.sqrt(bc3)).add_(eps)
    p.add_((m / bc1 + beta2 * v / bc2) / denom, alpha=-lr)
    if weight_decay != 0:
        p.div_(1 + weight_decay * lr)


@torch.no_grad()
def ademamix_step(p: Tensor, g: Tensor, m1: Tensor, m2: Tensor, nu: Tensor, step: int, lr: float, beta1: float, beta2: float,
                  beta3: float, alpha: float, eps: float, weight_decay: float = 0.0) -> None:
    """reference optim/ademamix.py:138-176: fast EMA bias-corrected, slow EMA (beta3) not."""
    bc1, bc2 = 1 - beta1**step, 1 - beta2**step
    if weight_decay != 0:
        g = g + weight_decay * p
    m1.mul_(beta1).add_(g, alpha=1 - beta1)
    nu.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    m2.mul_(beta3).add_(g, alpha=1 - beta3)
    denom = (nu.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m1 / bc1 + alpha * m2, denom, value=-lr)


@torch.no_grad()
def lars_step(p: Tensor, g: Tensor, buf: Optional[Tensor], lr: float, momentum: float = 0.0, dampening: float = 0.0,
              weight_decay: float = 0.0, nesterov: bool = False) -> Optional[Tensor]:
    """reference optim/lars.py:91-135. ``g`` is modified in place by the weight decay (as the reference does to p.grad);
    ``buf`` None on the first step with momentum -> a copy of the decayed gradient. Returns the momentum buffer."""
    p_norm = torch.norm(p)
    denom = torch.norm(g)
    if weight_decay != 0:
        g.add_(p, alpha=weight_decay)
        denom = denom + weight_decay * p_norm
    local_lr = 1.0 if (p_norm == 0 or denom == 0) else float(p_norm / denom)
    d = g
    if momentum != 0:
        if buf is None:
            buf = g.clone()
        else:
            buf.mul_(momentum).add_(g, alpha=1 - dampening)
        d = g.add(buf, alpha=momentum) if nesterov else buf
    p.add_(d, alpha=-lr * local_lr)
    return buf


@torch.no_grad()
def ralars_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, beta1: float, beta2: float, eps: float,
                weight_decay: float = 0.0, force_adaptive_momentum: bool = False,
                scale_clip: Tuple[float, float] = (0, 10)) -> float:
    """reference optim/ralars.py:56-140: RAdam update (rectified while the SMA length > 4) + LARS trust ratio.
    Returns the local learning-rate multiplier."""
    sma_inf = 2 / (1 - beta2) - 1
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1, bc2 = 1 - beta1**step, 1 - beta2**step
    sma_t = sma_inf - 2 * step * (1 - bc2) / bc2
    if sma_t > 4:
        r_t = math.sqrt((sma_t - 4) * (sma_t - 2) * sma_inf / ((sma_inf - 4) * (sma_inf - 2) * sma_t))
        update = r_t * (m / bc1) / ((v / bc2).sqrt() + eps)
    elif force_adaptive_momentum:
        update = (m / bc1) / ((v / bc2).sqrt() + eps)
    else:
        update = m / bc1
    if weight_decay != 0:
        update = update + weight_decay * p
    p_norm = p.pow(2).sum().sqrt()
    u_norm = update.pow(2).sum().sqrt()
    phi = p_norm.clamp(*scale_clip)
    local_lr = 1.0 if (phi == 0 or u_norm == 0) else float(phi / u_norm)
    p.add_(update, alpha=-lr * local_lr)
    return local_lr


@torch.no_grad()
def lookahead_sync(fast: Tensor, slow: Tensor, sync_rate: float) -> None:
    """reference optim/wrapper.py:122-135."""
    if sync_rate > 0:
        slow.add_(fast - slow, alpha=sync_rate)
    fast.copy_(slow)
