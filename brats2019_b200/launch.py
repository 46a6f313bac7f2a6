"""Training launcher for the hot path: what main.py:39-158 + the inner loop of
train.py:188-223 do, as one process per GPU instead of a single-process nn.DataParallel.

    python -m brats2019_b200.launch --steps 100                                   # one GPU, synthetic crops
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        -m brats2019_b200.launch --batchSize 16 --steps 1000 --models_path ./models --name run0

Same recipe as the reference: seed 1337 (main.py:44-48), UNet(depth 4, [1,2,2,4] / [1,1,1,1],
channels 16..128, 3 outputs) initialised like weight_init.py:22-27, criterion = mean(Dice_loss_joint,
BCE_Loss(bg_weight=1e-2)) (main.py:126-128, train.py:203-205), Adam(lr 2e-5, weight_decay 1e-6, amsgrad)
with StepLR(16000, 0.5) stepped once per batch (main.py:133-142, train.py:220-223).  `--batchSize` is the
GLOBAL batch as in main.py (DataParallel splits it across GPUs; here every rank takes batchSize / world).

Data: the reference's NIfTI readers (dataloader.py) are outside the hot path and need nibabel, which this
image does not have, so the launcher takes any iterable of `([image (B,4,D,H,W)], [target (B,3,D,H,W)])`
batches (the structure train.py:191-193 asserts) through `train(..., batches=...)`, and the command line
feeds it synthetic z-scored crops of the reference's patch size (144, 144, 128) (main.py:115).
Checkpoints are written in the reference's layout (checkpoint.save_reference_layout).
"""
from __future__ import annotations

import argparse
import contextlib
import math
import os
import random
import sys
import time

import torch
import torch.distributed as dist
import torch.nn as nn

from . import BCE_Loss, DEFAULT_CFG, Dice_loss_joint, UNet
from . import checkpoint as ckpt
from .graphs import GraphedTrainStep
from .optim import FusedAdam
from .parallel import DistributedUNet

PATCH = (144, 144, 128)        # main.py:115


def weight_init(m, generator=None):
    """weight_init.py:22-27 for the module types the UNet holds: Conv3d weights kaiming-normal
    (a = 1e-2, leaky_relu, fan_in), Conv3d biases N(0, 1); GroupNorm keeps its defaults."""
    if isinstance(m, nn.Conv3d):
        fan_in = m.weight.shape[1] * m.weight[0, 0].numel()
        std = math.sqrt(2.0 / (1.0 + 1e-2 ** 2)) / math.sqrt(fan_in)
        with torch.no_grad():
            m.weight.copy_(torch.randn(m.weight.shape, generator=generator) * std)
            if m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=generator))


def seed_everything(seed=1337):
    """main.py:44-48 (numpy is seeded too when present)."""
    random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    try:
        import numpy as np
        np.random.seed(seed)
    except ImportError:
        pass


def synthetic_batches(per_rank_batch, patch, steps, seed, device):
    """Z-scored random images (dataloader.py:141 normalises to zero mean / unit variance) and binary
    WT/TC/ET-style targets, generated on the host and copied like train.py:196-198 does."""
    g = torch.Generator().manual_seed(seed)
    for _ in range(steps):
        x = torch.randn(per_rank_batch, 4, *patch, generator=g).pin_memory()
        t = (torch.rand(per_rank_batch, 3, *patch, generator=g) > 0.7).float().pin_memory()
        yield [x.to(device, non_blocking=True)], [t.to(device, non_blocking=True)]


def build(device, cfg=DEFAULT_CFG, seed=1337, distributed=None):
    """main.py:56-61: model, init, wrapper.  Returns (net, criteria, process_group)."""
    with contextlib.redirect_stdout(sys.stderr):
        model = UNet(**cfg)
    g = torch.Generator().manual_seed(seed)          # every rank draws the same weights
    model.apply(lambda m: weight_init(m, g))
    model = model.to(device)
    distributed = dist.is_initialized() if distributed is None else distributed
    criteria = [Dice_loss_joint(index=0, priority=1), BCE_Loss(index=0, bg_weight=1e-2)]      # main.py:126-128
    net, group = model, None
    if distributed and dist.get_world_size() > 1:
        net = DistributedUNet(model)
        group = net.process_group
        for c in criteria:
            c.process_group = group
    return net, criteria, group


class MeanCriterion:
    """train.py:203-205: `loss_val = [c(output, target) for c in criterion]; loss = sum(loss_val) / len(loss_val)`.
    Keeps the per-criterion values of the last call (static tensors under a CUDA graph) for logging."""

    def __init__(self, criteria):
        self.criteria, self.vals = list(criteria), []

    def __call__(self, output, target):
        vals = [c(output, target) for c in self.criteria]
        self.vals = [v.detach() for v in vals]       # no reference to the autograd graph survives the step
        return sum(vals) / len(vals)


def train(net, criteria, batches, steps=None, lr=2e-5, weight_decay=1e-6, step_size=16000, gamma=0.5,
          log_every=10, log=print, state=None, fused=True, graph=True, eager_steps=2):
    """The batch loop of train.py:188-223: forward, mean of the criteria, backward, Adam step, scheduler step.

    fused=True (default): optim.FusedAdam - Adam(amsgrad, weight decay) and the StepLR schedule as one kernel over flat
    buffers; graph=True: from batch `eager_steps` on, the whole step (forward, criteria, backward, gradient all-reduce,
    optimizer) is ONE CUDA-graph replay per batch (a new graph is captured if the batch shape changes).  fused=False:
    the stock torch.optim.Adam(fused=True) + StepLR of round 1, eager.
    Returns (per-logged-step loss values per criterion as Python floats, TrainingState)."""
    if fused and graph and torch.cuda.current_stream() == torch.cuda.default_stream():
        # CUDA-graph capture cannot involve the legacy default stream (autograd's AccumulateGrad nodes run on the stream
        # they were created on): run the whole loop on a stream of our own, ordered after / before the caller's work
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            out = train(net, criteria, batches, steps, lr, weight_decay, step_size, gamma, log_every, log, state, fused, graph,
                        eager_steps)
        torch.cuda.current_stream().wait_stream(side)
        return out
    module = net.module if hasattr(net, "module") else net
    if fused:
        opt = FusedAdam(module.parameters(), lr=lr, weight_decay=weight_decay, amsgrad=True, lr_step_size=step_size,
                        lr_gamma=gamma, model=net)                                                                 # main.py:133-142
        sched = None
    else:
        opt = torch.optim.Adam(module.parameters(), lr=lr, weight_decay=weight_decay, amsgrad=True, fused=True)   # main.py:133-138
        sched = torch.optim.lr_scheduler.StepLR(opt, step_size=step_size, gamma=gamma)                             # main.py:139-142
    state = state if state is not None else ckpt.TrainingState()
    if state.optimizer_state is not None:
        opt.load_state_dict(state.optimizer_state)                                                            # train.py:83-84
    crit = MeanCriterion(criteria)
    history = []
    net.train()
    opt.zero_grad(set_to_none=True)
    gstep, gshape = None, None
    t0 = time.perf_counter()
    for idx, (data, target) in enumerate(batches):
        if steps is not None and idx >= steps:
            break
        assert isinstance(data, list) and isinstance(target, list)                                             # train.py:193
        if fused and graph and idx >= eager_steps:
            shape = (tuple(data[0].shape), tuple(target[0].shape))
            if gstep is None or shape != gshape:
                gstep, gshape = GraphedTrainStep(net, crit, opt, data[0], target[0], warmup=0), shape
            gstep(data[0], target[0])
        else:
            output = net(data)
            loss = crit(output, target)                                                                        # train.py:203-205
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            if sched is not None:
                sched.step()
        state.global_step += 1
        if log_every and (idx % log_every == 0):
            vals = [float(v.detach()) for v in crit.vals]         # one device->host read per logged step
            history.append(vals)
            lr_now = opt.param_groups[0]["lr"] * (gamma ** ((state.global_step - 1) // step_size) if fused else 1.0)
            log("step %d  loss %s  lr %.3g  %.2f s" % (state.global_step, ["%.5f" % v for v in vals], lr_now,
                                                     time.perf_counter() - t0))
    state.optimizer_state = opt.state_dict()                                                                   # train.py:315
    return history, state


def main(argv=None):
    ap = argparse.ArgumentParser(description="B200 launcher for the brats2019 ResUNet (replaces main.py)")
    ap.add_argument("--batchSize", type=int, default=2, help="GLOBAL batch (main.py --batchSize)")
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--patch", type=int, nargs=3, default=list(PATCH))
    ap.add_argument("--name", default="")
    ap.add_argument("--models_path", default="")
    ap.add_argument("--resume", default="", help="checkpoint in the reference layout (train.py:320-324)")
    ap.add_argument("--log-every", type=int, default=10)
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of one CUDA graph per step")
    ap.add_argument("--torch-adam", action="store_true", help="stock torch.optim.Adam(fused=True) + StepLR instead of FusedAdam")
    opt = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("brats2019_b200.launch needs a B200 (sm_100a); there is no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if opt.batchSize % world:
        raise RuntimeError("--batchSize %d is not divisible by the %d ranks" % (opt.batchSize, world))
    seed_everything(1337)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        net, criteria, _ = build(dev)
        state = None
        if opt.resume:
            state = ckpt.load_state_dict_into(net, opt.resume)
        batches = synthetic_batches(opt.batchSize // world, tuple(opt.patch), opt.steps, 100 + rank, dev)
        log = print if rank == 0 else (lambda *a, **k: None)
        _, state = train(net, criteria, batches, log_every=opt.log_every, log=log, state=state,
                         fused=not opt.torch_adam, graph=not opt.no_graph)
        torch.cuda.synchronize()
        if rank == 0 and opt.name and opt.models_path:
            d = os.path.join(opt.models_path, opt.name)
            os.makedirs(d, exist_ok=True)
            ckpt.save_reference_layout(os.path.join(d, opt.name + "last_model.pth"), net, state)            # train.py:313
            log("saved", os.path.join(d, opt.name + "last_model.pth"))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
