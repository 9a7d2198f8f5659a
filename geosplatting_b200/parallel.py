"""View sharding over ranks (one process per GPU) and the single per-step gradient all-reduce.

The reference is single-GPU and renders the views of a batch in a Python loop (rfstudio/model/geosplat.py:869-879,
loss = mean over views at rfstudio/trainer/geosplat_trainer.py:171-180).  A view's forward/backward touches only
the replicated Gaussian set and its own camera, so the path shards over views with NO data-path collective;
the only exchange is one sum all-reduce of the per-Gaussian (and env-map / exposure) gradients per step, on
one packed fp32 buffer (72 B per Gaussian) -- NCCL over NVLink on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, TypeVar

import torch
import torch.distributed as dist
from torch import Tensor

T = TypeVar("T")


def shard_views(items: Sequence[T], rank: int, world_size: int) -> List[T]:
    """Rank r renders views r, r+R, r+2R, ... of the global batch."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(items[rank::world_size])


def view_counts(n_views: int, world_size: int) -> List[int]:
    return [len(range(r, n_views, world_size)) for r in range(world_size)]


class GradientBucket:
    """One flat fp32 buffer holding every gradient that must be summed across ranks.

    pack() copies the tensors in (one fused foreach copy), all_reduce() issues ONE collective, unpack() hands
    back views.  The buffer is allocated once and reused every step."""

    def __init__(self, shapes: Sequence[torch.Size], device, dtype=torch.float32):
        self.shapes = [torch.Size(s) for s in shapes]
        self.sizes = [int(torch.Size(s).numel()) for s in self.shapes]
        self.flat = torch.zeros(sum(self.sizes), dtype=dtype, device=device)
        self.views: List[Tensor] = []
        o = 0
        for s, n in zip(self.shapes, self.sizes):
            self.views.append(self.flat[o:o + n].view(s))
            o += n

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def pack(self, tensors: Sequence[Optional[Tensor]]) -> None:
        assert len(tensors) == len(self.views)
        self.wait()
        src, dst = [], []
        for t, v in zip(tensors, self.views):
            if t is None:
                v.zero_()
            else:
                assert t.shape == v.shape, (t.shape, v.shape)
                src.append(t.detach())
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)

    def all_reduce(self, group=None, average_over: Optional[int] = None, async_op: bool = False) -> None:
        """ONE sum all-reduce of the whole bucket.  With async_op=True the collective runs on NCCL's stream and
        overlaps the next views' kernels; call wait() before the bucket is read or packed again."""
        if average_over:
            self.flat.mul_(1.0 / average_over)   # pre-scale: sum of (g / n) == mean
        self._work = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            if async_op:
                self._work = work

    def wait(self) -> None:
        work = getattr(self, "_work", None)
        if work is not None:
            work.wait()
            self._work = None

    def unpack(self) -> List[Tensor]:
        self.wait()
        return self.views


def common_flat(tensors: Sequence[Optional[Tensor]]) -> Optional[Tensor]:
    """The contiguous 1-D tensor that covers every gradient when they are all views of ONE buffer (the batched backward of
    fused.splat_views writes them into a single flat buffer: [env | quats | ks | means | scales | logits | normals | kd |
    exposure]), else None.  Bytes between the views (alignment padding, per-view slots) are covered too."""
    ts = [t for t in tensors if t is not None]
    if not ts:
        return None
    st = ts[0].untyped_storage()
    for t in ts:
        if t.untyped_storage().data_ptr() != st.data_ptr() or not t.is_contiguous() or t.dtype != ts[0].dtype:
            return None
    lo = min(t.storage_offset() for t in ts)
    hi = max(t.storage_offset() + t.numel() for t in ts)
    return torch.empty(0, dtype=ts[0].dtype, device=ts[0].device).set_(st, lo, (hi - lo,))


class FlatAllReduce:
    """ONE sum all-reduce of a batch's gradients IN PLACE, without packing: `tensors` are what torch.autograd.grad
    returned for fused.splat_views.  Falls back to a packed bucket when the gradients do not share a buffer."""

    def __init__(self, tensors: Sequence[Optional[Tensor]], group=None, async_op: bool = False):
        self.flat = common_flat(tensors)
        self.packed = self.flat is None
        self._work = None
        self._bucket = None
        if self.packed:
            ts = [t for t in tensors if t is not None]
            self._bucket = GradientBucket([t.shape for t in ts], ts[0].device, ts[0].dtype)
            self._bucket.pack(ts)
            self.flat = self._bucket.flat
            self._targets = ts
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
            self._work = work if async_op else None

    @property
    def nbytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def wait(self) -> None:
        if self._work is not None:
            self._work.wait()
            self._work = None
        if self.packed and self._bucket is not None:
            for t, v in zip(self._targets, self._bucket.views):
                t.copy_(v)
            self._bucket = None


def allreduce_gradients(tensors: Sequence[Optional[Tensor]], bucket: Optional[GradientBucket] = None, group=None,
                        average_over: Optional[int] = None) -> List[Tensor]:
    """Sum `tensors` across ranks with a single collective; returns views into the bucket."""
    if bucket is None:
        if any(t is None for t in tensors):
            raise ValueError("allreduce_gradients: pass a GradientBucket when some gradients are None")
        ref = tensors[0]
        bucket = GradientBucket([t.shape for t in tensors], ref.device, ref.dtype)
    bucket.pack(tensors)
    bucket.all_reduce(group, average_over)
    return bucket.unpack()
