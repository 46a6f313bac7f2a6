"""Drop-in mirror of the reference's criteria on the hot path (loss.py:64-79, 98-122).

`Dice_loss_joint` and `BCE_Loss` keep the reference's constructors and the
`criterion(output_list, target_list)` call shape (train.py:203-205) and run as hand-written CUDA
reductions (b200_dice_*, b200_bce_*).  The
six per-channel sums are produced without the epsilons so that a data-parallel run can
all-reduce them before the loss is formed: set `Dice_loss_joint.process_group`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _as_target(t):
    """fp32 targets as the reference passes them (soft values allowed, loss.py:105-111), or binary masks staged as
    uint8 / bool, which the kernels convert on the fly (a quarter of the host->device bytes)."""
    t = t.contiguous()
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    return t if t.dtype in (torch.float32, torch.uint8) else t.float()


class _DiceFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, target, priority, group):
        probs = probs.contiguous()
        target = _as_target(target)
        sums = ops.dice_sums(probs, target)
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)      # 8 floats
        loss = ops.dice_loss(sums, probs.shape[1], priority)
        ctx.save_for_backward(probs, target, sums)
        ctx.priority = priority
        return loss.reshape(())

    @staticmethod
    def backward(ctx, gout):
        probs, target, sums = ctx.saved_tensors
        gout = gout.reshape(1).float().contiguous()
        gp = ops.dice_backward(probs, target, sums, gout, ctx.priority)
        return gp, None, None, None


class Dice_loss_joint(nn.Module):
    """loss.py:98-122: priority * (1 - mean_c 2(sum p g + 1e-6) / (sum (p^2 + g) + 2e-6)),
    sums over batch and space."""

    process_group = None     # torch.distributed group whose ranks share one global batch

    def __init__(self, index=0, priority=1):
        super(Dice_loss_joint, self).__init__()
        self.index = index
        self.priority = priority

    def forward(self, x, y):
        pred, gt = x[self.index], y[self.index]
        assert (pred.shape == gt.shape)
        if not pred.is_cuda:
            raise RuntimeError("brats2019_b200.Dice_loss_joint runs on CUDA only (no CPU fallback)")
        if pred.shape[1] > 4:
            raise RuntimeError("brats2019_b200.Dice_loss_joint supports at most 4 channels")
        with torch.cuda.device(pred.device):
            return _DiceFunction.apply(pred.float(), gt, float(self.priority), self.process_group)


class _BCEFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, probs, target, bg_weight, group):
        probs = probs.contiguous()
        target = _as_target(target)
        s = ops.bce_sum(probs, target, bg_weight)
        numel = probs.numel()
        if group is not None:
            import torch.distributed as dist
            dist.all_reduce(s, op=dist.ReduceOp.SUM, group=group)
            numel *= dist.get_world_size(group)          # the mean runs over the global batch
        ctx.save_for_backward(probs, target)
        ctx.bg_weight, ctx.numel = bg_weight, numel
        return ops.bce_loss(s, numel).reshape(())

    @staticmethod
    def backward(ctx, gout):
        probs, target = ctx.saved_tensors
        gout = gout.reshape(1).float().contiguous()
        return ops.bce_backward(probs, target, gout, ctx.bg_weight, ctx.numel), None, None, None


class BCE_Loss(nn.Module):
    """loss.py:64-79 — the other half of the trainer's criterion (main.py:127):
    -mean(g log(p+1e-6) + bg_weight (1-g) log(1+1e-6-p)), as CUDA reductions (b200_bce_*)."""

    process_group = None

    def __init__(self, index=0, bg_weight=1):
        super(BCE_Loss, self).__init__()
        self.label_index = index
        self.bg_weight = bg_weight

    def forward(self, x, y):
        assert (x[self.label_index].shape == y[self.label_index].shape)
        pred = x[self.label_index]
        gt = y[self.label_index]
        if not pred.is_cuda:
            raise RuntimeError("brats2019_b200.BCE_Loss runs on CUDA only (no CPU fallback)")
        with torch.cuda.device(pred.device):
            return _BCEFunction.apply(pred.float(), gt, float(self.bg_weight), self.process_group)
