"""Data-parallel plumbing for the hot path: one process per GPU, one flat gradient bucket, one NCCL all-reduce per
optimizer step over NVLink/NVSwitch.

The reference has no distributed code at all (a single ``gpu`` argument in
holocron/trainer/core.py:52, 90-104). The path shards by batch: every rank runs the same fused kernels on its shard,
BatchNorm statistics stay per-GPU (the reference uses plain ``nn.BatchNorm2d``), and the only exchange is the mean of the
parameter gradients. ``torch.distributed`` (backend ``nccl`` on GPUs, ``gloo`` in the CPU tests) is the plumbing.

Design: all ``.grad`` tensors are views into ONE contiguous fp32 bucket (allocated once, laid out in reverse
registration order = roughly the order backward produces them), so
  * backward accumulates straight into the bucket (no per-step flatten copy),
  * the all-reduce is a single collective on the bucket (size ~100 MB for RepVGG: latency-, not bandwidth-bound),
  * the fused optimizer kernels read the reduced gradients in place through the same pointers.
"""
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist
from torch import Tensor, nn


class GradBucket:
    """Flat fp32 gradient storage whose slices are installed as the parameters' ``.grad``."""

    def __init__(self, params: Iterable[nn.Parameter], direct: bool = True) -> None:
        """``direct``: allow the backward kernels of this package (weight-gradient reduction, BatchNorm parameter
        gradients) to ADD their results straight into the bucket views and hand ``None`` to autograd for those
        parameters, instead of materialising a gradient tensor that AccumulateGrad adds with one more element-wise kernel
        per parameter and step (207 launches per RepVGG-A0 step). Tensor hooks registered on such parameters do not fire."""
        self.params: List[nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        # 64-element alignment keeps every view 256-byte aligned for the vectorised optimizer kernels
        offsets, total = [], 0
        for p in reversed(self.params):
            offsets.append(total)
            total += (p.numel() + 63) // 64 * 64
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.views: List[Tensor] = []
        for p, off in zip(reversed(self.params), offsets):
            chunk = self.flat[off:off + p.numel()]
            if p.ndim == 4 and p.is_contiguous(memory_format=torch.channels_last) and not p.is_contiguous():
                n, c, h, w = p.shape
                view = chunk.view(n, h, w, c).permute(0, 3, 1, 2)
            else:
                view = chunk.view(p.shape) if p.is_contiguous() else chunk.view(-1)[:p.numel()].view(p.shape)
            self.views.append(view)
            p.grad = view
            p._hb_direct_grad = bool(direct)

    def zero_(self) -> None:
        """Replaces ``optimizer.zero_grad()``: one memset, gradients stay bound to the bucket."""
        self.flat.zero_()
        for p, v in zip(reversed(self.params), self.views):
            if p.grad is not v:
                p.grad = v

    def all_reduce_mean(self, group: Optional[dist.ProcessGroup] = None, async_op: bool = False):
        """Averages the bucket across ranks (sum all-reduce on the pre-scaled bucket)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        self.flat.mul_(1.0 / dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def broadcast_parameters(module: nn.Module, src: int = 0, group: Optional[dist.ProcessGroup] = None) -> None:
    """Makes every rank start from rank ``src``'s parameters and buffers (replicas are kept identical afterwards by the
    gradient all-reduce alone)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=src, group=group)


def shard_batch(global_batch: int, rank: int, world_size: int) -> range:
    """Contiguous batch shard of rank ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))
