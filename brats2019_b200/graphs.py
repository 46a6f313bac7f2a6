"""CUDA-graph capture of the hot loop.  A training step is ~370 kernel launches of a few
microseconds to ~0.2 ms each; issued one by one from Python they are CPU-bound.  Every kernel of
this package is stream-ordered, allocation-free after warm-up and never synchronises, so the whole
step (train.py:201-221: forward, criterion, backward, optimizer.step) replays as ONE graph launch.
"""
from __future__ import annotations

import torch


class GraphedTrainStep:
    """step = zero_grad; loss = criterion(model([x]), [t]); loss.backward(); optimizer.step()

    `optimizer` must be capturable (e.g. torch.optim.Adam(..., capturable=True)).  Inputs are
    copied into static buffers; `__call__` returns the static 0-dim loss tensor of the replay.
    """

    def __init__(self, model, criterion, optimizer, x, t, warmup=3):
        self.model, self.criterion, self.optimizer = model, criterion, optimizer
        self.x = x.detach().clone()
        self.t = t.detach().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager()
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        with torch.cuda.graph(self.graph):
            self.loss = self._eager(zero=False)

    def _eager(self, zero=True):
        if zero:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.criterion(self.model([self.x]), [self.t])
        loss.backward()
        self.optimizer.step()
        return loss

    def __call__(self, x=None, t=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if t is not None:
            self.t.copy_(t, non_blocking=True)
        self._replay()
        return self.loss

    def _replay(self):
        sync_lr = getattr(self.optimizer, "sync_lr", None)
        if sync_lr is not None:
            sync_lr()                    # a host LR scheduler may have rewritten param_groups[0]['lr']
        self.graph.replay()
        # the replayed optimizer step rewrote the weights without touching their version counters: the next EAGER
        # forward (an eval pass between training phases) must re-pack its bf16 weight images
        m = self.model.module if hasattr(self.model, "module") else self.model
        if hasattr(m, "invalidate_packed_weights"):
            m.invalidate_packed_weights()

    # -- host-input pipeline: the H2D copy of step i+1 overlaps the compute of step i -------------
    def prefetch(self, x_host, t_host):
        """Start copying the NEXT step's (pinned) host batch on a copy stream into a staging slot."""
        if not hasattr(self, "_stage"):
            self._copy_stream = torch.cuda.Stream()
            self._stage = [(torch.empty_like(self.x), torch.empty_like(self.t)) for _ in range(2)]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            self._slot_in, self._slot_out = 0, 0
            for e in self._free:
                e.record()
        k = self._slot_in
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(self._free[k])          # the step that read this slot has consumed it
            self._stage[k][0].copy_(x_host, non_blocking=True)
            self._stage[k][1].copy_(t_host, non_blocking=True)
            self._ready[k].record()
        self._slot_in = k ^ 1

    def step_prefetched(self):
        """Run one step on the oldest prefetched batch; returns the static loss tensor."""
        k = self._slot_out
        cur = torch.cuda.current_stream()
        cur.wait_event(self._ready[k])
        self.x.copy_(self._stage[k][0], non_blocking=True)       # device-to-device, ~40 us
        self.t.copy_(self._stage[k][1], non_blocking=True)
        self._free[k].record()
        self._slot_out = k ^ 1
        self._replay()
        return self.loss


class GraphedForward:
    """probs = model([x])[0] under no_grad (Trainer.predict, train.py:129-143) as one graph launch."""

    def __init__(self, model, x, warmup=2):
        self.model = model
        self.x = x.detach().clone()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):
                model([self.x])
        cur.wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.probs = model([self.x])[0]

    def __call__(self, x=None):
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.probs
