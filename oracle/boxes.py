"""Oracle restatements of holocron.ops.boxes (reference: holocron/ops/boxes.py; IoU from torchvision.ops.boxes)."""
import math

import torch
from torch import Tensor


def _area(b: Tensor) -> Tensor:
    return (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])


def _inter_union(b1: Tensor, b2: Tensor):
    # torchvision.ops.boxes._box_inter_union: same operation order so exact-equality vectors hold
    a1, a2 = _area(b1), _area(b2)
    x0 = torch.max(b1[:, None, 0], b2[None, :, 0])
    y0 = torch.max(b1[:, None, 1], b2[None, :, 1])
    x1 = torch.min(b1[:, None, 2], b2[None, :, 2])
    y1 = torch.min(b1[:, None, 3], b2[None, :, 3])
    inter = (x1 - x0).clamp(min=0) * (y1 - y0).clamp(min=0)
    union = a1[:, None] + a2[None, :] - inter
    return inter, union


def box_iou(b1: Tensor, b2: Tensor) -> Tensor:
    inter, union = _inter_union(b1, b2)
    return inter / union


def box_giou(b1: Tensor, b2: Tensor) -> Tensor:
    """IoU - (|C| - |A u B|) / |C| with C the enclosing box — reference ops/boxes.py:33-66."""
    if (b1[:, 2:] < b1[:, :2]).any() or (b2[:, 2:] < b2[:, :2]).any():
        raise AssertionError("Incorrect coordinate format")
    inter, union = _inter_union(b1, b2)
    ew = (torch.max(b1[:, None, 2], b2[None, :, 2]) - torch.min(b1[:, None, 0], b2[None, :, 0])).clamp(min=0)
    eh = (torch.max(b1[:, None, 3], b2[None, :, 3]) - torch.min(b1[:, None, 1], b2[None, :, 1])).clamp(min=0)
    enclosing = ew * eh
    return inter / union - (enclosing - union) / enclosing


def iou_penalty(b1: Tensor, b2: Tensor) -> Tensor:
    """rho^2 / c^2: squared centre distance over squared enclosing-box diagonal — reference ops/boxes.py:69-103.
    Always fp32 (the reference allocates its intermediates with torch.zeros without dtype)."""
    b1, b2 = b1.float(), b2.float()
    dw = torch.max(b1[:, None, 2], b2[None, :, 2]) - torch.min(b1[:, None, 0], b2[None, :, 0])
    dh = torch.max(b1[:, None, 3], b2[None, :, 3]) - torch.min(b1[:, None, 1], b2[None, :, 1])
    c2 = dw**2 + dh**2
    cx = (b1[:, 0] + b1[:, 2])[:, None] - (b2[:, 0] + b2[:, 2])[None, :]
    cy = (b1[:, 1] + b1[:, 3])[:, None] - (b2[:, 1] + b2[:, 3])[None, :]
    return (cx**2 + cy**2) / 4 / c2


def diou_loss(b1: Tensor, b2: Tensor) -> Tensor:
    """1 - IoU + rho^2/c^2 — reference ops/boxes.py:106-130."""
    return 1 - box_iou(b1, b2) + iou_penalty(b1, b2)


def aspect_ratio(b: Tensor) -> Tensor:
    """atan(w / h) — reference ops/boxes.py:133-142."""
    return torch.atan((b[:, 2] - b[:, 0]) / (b[:, 3] - b[:, 1]))


def aspect_ratio_consistency(b1: Tensor, b2: Tensor) -> Tensor:
    """4/pi^2 (atan(w1/h1) - atan(w2/h2))^2 — reference ops/boxes.py:145-159."""
    d = aspect_ratio(b1)[:, None] - aspect_ratio(b2)[None, :]
    return d**2 * (4 / math.pi**2)


def ciou_loss(b1: Tensor, b2: Tensor) -> Tensor:
    """Reference ops/boxes.py:162-211. The alpha*v term is added to a masked *copy* there (line 209), so the
    returned value is exactly the DIoU loss; reproduced as such."""
    return diou_loss(b1, b2)
