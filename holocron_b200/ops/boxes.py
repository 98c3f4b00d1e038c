"""Box operators on the fused pairwise CUDA kernels — API mirror of holocron/ops/boxes.py."""
from typing import Tuple

import torch
from torch import Tensor

from .._lib import check, lib, ptr, require_cuda, stream_ptr

__all__ = ["box_giou", "ciou_loss", "diou_loss"]

_IOU, _GIOU, _PENALTY, _DIOU_LOSS, _ARC = range(5)


def _prep(boxes: Tensor) -> Tensor:
    if boxes.ndim != 2 or boxes.shape[1] != 4:
        raise ValueError("boxes are expected as (num_boxes, 4) tensors in xyxy format")
    # box arithmetic always runs in fp32 (the reference's intermediates are fp32 whatever the input dtype)
    out = boxes.float().contiguous()
    # the kernels read a box as one 16-byte load
    return out if out.data_ptr() % 16 == 0 else out.clone()


class _PairwiseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, boxes1: Tensor, boxes2: Tensor, mode: int) -> Tensor:
        require_cuda(boxes1, boxes2)
        b1, b2 = _prep(boxes1), _prep(boxes2)
        m, n = b1.shape[0], b2.shape[0]
        out = torch.empty((m, n), device=b1.device, dtype=torch.float32)
        check(lib().hb_box_pairwise(ptr(b1), ptr(b2), ptr(out), m, n, mode, stream_ptr()), "hb_box_pairwise")
        ctx.save_for_backward(b1, b2)
        ctx.mode = mode
        ctx.in_dtypes = (boxes1.dtype, boxes2.dtype)
        return out

    @staticmethod
    def backward(ctx, gout: Tensor):
        b1, b2 = ctx.saved_tensors
        if ctx.mode == _ARC:
            raise NotImplementedError("aspect_ratio_consistency has no backward kernel (unused by the reference losses)")
        m, n = b1.shape[0], b2.shape[0]
        g = gout.float().contiguous()
        g1 = torch.empty_like(b1) if ctx.needs_input_grad[0] else None
        g2 = torch.empty_like(b2) if ctx.needs_input_grad[1] else None
        check(lib().hb_box_pairwise_bwd(ptr(b1), ptr(b2), ptr(g), ptr(g1), ptr(g2), m, n, ctx.mode, stream_ptr()),
              "hb_box_pairwise_bwd")
        d1, d2 = ctx.in_dtypes
        return (None if g1 is None else g1.to(d1)), (None if g2 is None else g2.to(d2)), None


def box_iou(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """Pairwise IoU (what the reference takes from torchvision.ops.boxes.box_iou, boxes.py:11)."""
    return _PairwiseFn.apply(boxes1, boxes2, _IOU)


def _box_iou(boxes1: Tensor, boxes2: Tensor) -> Tuple[Tensor, Tensor]:
    """(iou, union) like reference boxes.py:16-30; the union is recovered from the two kernel outputs."""
    iou = box_iou(boxes1, boxes2)
    b1, b2 = _prep(boxes1), _prep(boxes2)
    area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
    area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
    union = (area1[:, None] + area2) / (1 + iou)
    return iou, union


def box_giou(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """Generalized IoU — mirrors reference boxes.py:33-66, including the ``AssertionError`` on degenerate boxes
    (which, as in the reference, needs one device synchronisation)."""
    require_cuda(boxes1, boxes2)
    b1, b2 = _prep(boxes1), _prep(boxes2)
    flag = torch.zeros(1, device=b1.device, dtype=torch.int32)
    check(lib().hb_box_degenerate(ptr(b1), b1.shape[0], ptr(flag), stream_ptr()), "hb_box_degenerate")
    check(lib().hb_box_degenerate(ptr(b2), b2.shape[0], ptr(flag), stream_ptr()), "hb_box_degenerate")
    if int(flag.item()) != 0:
        raise AssertionError("Incorrect coordinate format")
    return _PairwiseFn.apply(boxes1, boxes2, _GIOU)


def iou_penalty(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """DIoU penalty: squared centre distance over squared enclosing-box diagonal (reference boxes.py:69-103)."""
    return _PairwiseFn.apply(boxes1, boxes2, _PENALTY)


def diou_loss(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """Distance-IoU loss ``1 - IoU + rho^2/c^2`` (reference boxes.py:106-130) in one fused launch."""
    return _PairwiseFn.apply(boxes1, boxes2, _DIOU_LOSS)


def aspect_ratio(boxes: Tensor) -> Tensor:
    """atan(w / h) (reference boxes.py:133-142). N-sized host-side helper kept in torch."""
    require_cuda(boxes)
    return torch.atan((boxes[:, 2] - boxes[:, 0]) / (boxes[:, 3] - boxes[:, 1]))


def aspect_ratio_consistency(boxes1: Tensor, boxes2: Tensor) -> Tensor:
    """``4/pi^2 (atan(w1/h1) - atan(w2/h2))^2`` (reference boxes.py:145-159)."""
    return _PairwiseFn.apply(boxes1, boxes2, _ARC)


def ciou_loss(boxes1: Tensor, boxes2: Tensor, paper_correct: bool = False) -> Tensor:
    """Complete-IoU loss as the REFERENCE computes it (boxes.py:162-211): its ``alpha * v`` term is accumulated
    into a boolean-mask copy and discarded, so the result is bit-for-bit the DIoU loss. That behaviour is the
    default here; ``paper_correct=True`` (not in the reference API) adds ``alpha * v`` with
    ``alpha = v / (1 - IoU + v)`` as in the paper."""
    loss = diou_loss(boxes1, boxes2)
    if not paper_correct:
        return loss
    v = aspect_ratio_consistency(boxes1, boxes2).detach()
    iou = box_iou(boxes1, boxes2)
    alpha = (v / (1 - iou + v).clamp_min(1e-12)).detach()
    return loss + alpha * v
