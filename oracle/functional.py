"""Oracle restatements of holocron.nn.functional (reference: holocron/nn/functional.py)."""
import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor


def hard_mish(x: Tensor) -> Tensor:
    """x/2 * min(max(x+2, 0), 2)  — reference nn/functional.py:30-41."""
    gate = torch.clamp(x + 2, min=0, max=2)  # clamp's inclusive sub-gradient at the kinks is part of the semantics
    return x * 0.5 * gate


def nl_relu(x: Tensor, beta: float = 1.0) -> Tensor:
    """log(1 + beta * max(x, 0))  — reference nn/functional.py:44-56."""
    return torch.log(1 + beta * torch.relu(x))  # relu (sub-gradient 0 at x = 0), as in the reference


def _rows(x: Tensor) -> Tensor:
    """(N, K, *spatial) -> (N*prod(spatial), K) with the reference's (n, spatial) flattening order."""
    k = x.shape[1]
    return x.reshape(x.shape[0], k, -1).permute(0, 2, 1).reshape(-1, k)


def focal_loss(x: Tensor, target: Tensor, weight: Optional[Tensor] = None, ignore_index: int = -100,
               reduction: str = "mean", gamma: float = 2.0) -> Tensor:
    """-(w_t) (1 - p_t)^gamma log p_t  — reference nn/functional.py:59-113.

    Quirks kept: the class weight scales log p_t only; ignore_index is honoured only inside [0, K);
    'mean' divides by the number of non-ignored positions; 'none' is reshaped like the target.
    """
    k = x.shape[1]
    logp = F.log_softmax(_rows(x), dim=1)
    t = target.reshape(-1)
    logpt = logp[torch.arange(t.numel()), t]
    pt = logpt.exp()
    if weight is not None:
        logpt = weight.to(x.dtype)[t] * logpt
    loss = -((1 - pt) ** gamma) * logpt
    keep = torch.ones_like(t, dtype=torch.bool)
    if 0 <= ignore_index < k:
        keep = t != ignore_index
    if reduction == "sum":
        return loss[keep].sum()
    if reduction == "mean":
        return loss[keep].mean()
    return loss.reshape(target.shape)


def poly_loss(x: Tensor, target: Tensor, eps: float = 2.0, weight: Optional[Tensor] = None,
              ignore_index: int = -100, reduction: str = "mean") -> Tensor:
    """Poly-1 loss: -log p_t + eps (1 - p_t)  — reference nn/functional.py:540-613 (hard and soft targets)."""
    k = x.shape[1]
    hard = target.ndim == x.ndim - 1
    if hard:
        if target.dtype != torch.long:
            raise TypeError("target dtype is expected to be torch.int64")
        logp = F.log_softmax(_rows(x), dim=1)
        t = target.reshape(-1)
        z = logp[torch.arange(t.numel()), t]
        loss = -z + eps * (1 - z.exp())
        if weight is not None:
            loss = weight.to(x.dtype)[t] * loss
        keep = torch.ones_like(t, dtype=torch.bool)
        if 0 <= ignore_index < k:
            keep = t != ignore_index
        if reduction == "sum":
            return loss[keep].sum()
        if reduction == "mean":
            return loss[keep].mean()
        return loss  # flat, as in the reference
    if target.ndim != x.ndim or target.shape[:2] != x.shape[:2]:
        raise ValueError("invalid target shape")
    z = F.log_softmax(x, dim=1) * target
    loss = -z + eps * (1 - z.exp())
    if weight is not None:
        # reference: weight.reshape(1, -1) * loss -> broadcasts against the LAST dim; only valid for (N, K) inputs
        loss = weight.to(x.dtype).reshape(1, -1) * loss
    cls = [c for c in range(k) if not (0 <= ignore_index < k and c == ignore_index)]
    sel = loss[:, cls]
    if reduction == "sum":
        return sel.sum()
    if reduction == "mean":
        return sel.sum(1).mean()
    return sel.sum(1)


def dice_loss(x: Tensor, target: Tensor, weight: Optional[Tensor] = None, gamma: float = 1.0,
              eps: float = 1e-8) -> Tensor:
    """1 - (1 + 1/gamma) mean_k[(gamma sum(x t) + eps) / (sum(x + gamma t) + eps)]  — reference nn/functional.py:503-537."""
    k = x.shape[1]
    xs = x.transpose(0, 1).reshape(k, -1)
    ts = target.transpose(0, 1).reshape(k, -1)
    inter = gamma * (xs * ts).sum(1)
    card = (xs + gamma * ts).sum(1)
    coeff = (inter + eps) / (card + eps)
    if weight is None:
        return 1 - (1 + 1 / gamma) * coeff.mean()
    w = weight.to(x.dtype)
    return 1 - (1 + 1 / gamma) * (w * coeff).sum() / w.sum()


def _patches(x: Tensor, kh: int, kw: int, stride: int, padding: int, dilation: int) -> Tensor:
    """(N, C, H, W) -> (N, L, C*kh*kw) sliding windows, channel-major inside a window (the im2col order of F.unfold)."""
    return F.unfold(x, (kh, kw), dilation=dilation, padding=padding, stride=stride).transpose(1, 2)


def _out_hw(h: int, w: int, kh: int, kw: int, stride: int, padding: int, dilation: int):
    ho = (h + 2 * padding - dilation * (kh - 1) - 1) // stride + 1
    wo = (w + 2 * padding - dilation * (kw - 1) - 1) // stride + 1
    return ho, wo


def norm_conv2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0,
                dilation: int = 1, eps: float = 1e-14) -> Tensor:
    """Each im2col patch (whole Cin*kh*kw vector, zero padding included) is standardised with its biased
    variance (+eps inside the rsqrt) and then correlated with the filters — reference nn/functional.py:322-413.
    ``groups`` is ignored by the reference and therefore absent here."""
    co, _, kh, kw = weight.shape
    p = _patches(x, kh, kw, stride, padding, dilation)
    mu = p.mean(-1, keepdim=True)
    var = ((p - mu) ** 2).mean(-1, keepdim=True)
    pn = (p - mu) * torch.rsqrt(var + eps)
    out = pn @ weight.reshape(co, -1).t()
    if bias is not None:
        out = out + bias
    ho, wo = _out_hw(x.shape[2], x.shape[3], kh, kw, stride, padding, dilation)
    return out.transpose(1, 2).reshape(x.shape[0], co, ho, wo)


def add2d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, padding: int = 0,
          dilation: int = 1, normalize_slices: bool = False, eps: float = 1e-14) -> Tensor:
    """AdderNet response: -sum_k |patch_k - w_{c,k}|  — reference nn/functional.py:416-462."""
    co, _, kh, kw = weight.shape
    p = _patches(x, kh, kw, stride, padding, dilation)
    if normalize_slices:
        mu = p.mean(-1, keepdim=True)
        var = ((p - mu) ** 2).mean(-1, keepdim=True)
        p = (p - mu) * torch.rsqrt(var + eps)
    wf = weight.reshape(co, -1)
    out = torch.stack([-(p - wf[c]).abs().sum(-1) for c in range(co)], dim=-1)
    if bias is not None:
        out = out + bias
    ho, wo = _out_hw(x.shape[2], x.shape[3], kh, kw, stride, padding, dilation)
    return out.transpose(1, 2).reshape(x.shape[0], co, ho, wo)


def dropblock2d_with_noise(x: Tensor, noise: Tensor, drop_prob: float, block_size: int) -> Tensor:
    """DropBlock given the uniform noise tensor (N, H, W) the reference would have drawn
    (reference nn/functional.py:465-500): seeds = noise <= drop_prob / block_size**2, dilated by a
    block_size max-pool, shared across channels, rescaled by numel/kept."""
    gamma = drop_prob / block_size**2
    seeds = (noise <= gamma).to(x.dtype)
    mask = 1 - F.max_pool2d(seeds, (block_size, block_size), stride=1, padding=block_size // 2)
    kept = mask.sum()
    out = x * mask.unsqueeze(1)
    if kept > 0:
        out = out * (mask.numel() / kept)
    return out
