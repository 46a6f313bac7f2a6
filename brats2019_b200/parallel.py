"""Data parallelism for the ResUNet hot path: one process per GPU, weights replicated and
resident, batch split by volume.  Replaces the reference's single-process
`nn.DataParallel` (main.py:61), which re-broadcasts the weights and gathers outputs every step.

Per step there are exactly two collectives, both NCCL all-reduces over NVLink:
  * 8 floats inside the Dice forward (the loss sums over the global batch, loss.py:114-115);
  * the ~4.5 M live gradient elements, as views of one flat fp32 buffer, issued bucket by
    bucket from inside backward (Engine.backward -> BucketedAllReduce.mark) so the transfer of
    the deep levels overlaps the level-0 backward, which is 44 % of the FLOPs.
Gradients are SUMMED (not averaged): with the Dice sums all-reduced, each rank's backward
yields the derivative of the *global* loss w.r.t. the parameters through its own samples, and
their sum is the single-process gradient on the concatenated batch - the reference's
DataParallel semantics.  The 7 dead parameters never get gradients and are never reduced.
"""
from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn as nn

from .engine import GradStore


class BucketedAllReduce(GradStore):
    """GradStore whose tensors are views of one flat buffer, all-reduced in buckets."""

    def __init__(self, flat, group, min_bucket_elems=1 << 20):
        super().__init__()
        self.flat, self.group, self.min_bucket = flat, group, min_bucket_elems
        self.cursor = 0
        self.start = 0
        self.handles = []
        self.buckets = []          # (start, end) element ranges, for tests / introspection

    def new(self, name, like):
        n = like.numel()
        if self.cursor + n > self.flat.numel():
            raise RuntimeError("gradient flat buffer too small")
        t = self.flat[self.cursor:self.cursor + n].view(like.shape)
        self.cursor += (n + 3) // 4 * 4        # keep every view 16-byte aligned
        self.grads[name] = t
        return t

    def _launch(self):
        if self.cursor > self.start:
            seg = self.flat[self.start:self.cursor]
            self.handles.append(dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            self.buckets.append((self.start, self.cursor))
            self.start = self.cursor

    def mark(self):
        if self.cursor - self.start >= self.min_bucket:
            self._launch()

    def finish(self):
        self._launch()
        for h in self.handles:
            h.wait()                # stream-orders the consumer after the collective; no host sync on CUDA
        self.handles = []


def flat_grad_elems(module):
    return sum((p.numel() + 3) // 4 * 4 for _, p in module.live_parameters())


class DistributedUNet(nn.Module):
    """Wrapper exposing `.module` like nn.DataParallel (train.py:290, export_onnx.py:22 reach
    through it).  Call `criterion.process_group = wrapper.process_group` on Dice_loss_joint."""

    def __init__(self, module, process_group=None, min_bucket_elems=1 << 20, broadcast=True):
        super().__init__()
        self.module = module
        self.process_group = process_group if process_group is not None else dist.group.WORLD
        self.world_size = dist.get_world_size(self.process_group)
        self.min_bucket_elems = min_bucket_elems
        if broadcast:
            for p in module.parameters():
                dist.broadcast(p.data, src=dist.get_global_rank(self.process_group, 0), group=self.process_group)
            module.invalidate_packed_weights()       # `.data` writes do not bump version counters
        module._grad_store_factory = self._make_store
        self.flat_provider = None      # optim.FusedAdam.bind(): persistent flat gradient buffer shared with the optimizer
        self.last_buckets = []

    def _make_store(self):
        dev = next(self.module.parameters()).device
        # A fresh flat buffer per step (18 MB from the caching allocator): autograd may adopt the
        # returned views as `.grad`, so the storage must not be reused by the next backward.
        flat = self.flat_provider() if self.flat_provider is not None else None
        if flat is None:
            flat = torch.empty(flat_grad_elems(self.module), dtype=torch.float32, device=dev)
        store = BucketedAllReduce(flat, self.process_group, self.min_bucket_elems)
        self.last_buckets = store.buckets
        return store

    def forward(self, x):
        return self.module(x)


def shard_volumes(n_volumes, rank, world_size):
    """Inference partition: contiguous shard of the volume list, no collective."""
    per = (n_volumes + world_size - 1) // world_size
    lo = min(rank * per, n_volumes)
    return range(lo, min(lo + per, n_volumes))
