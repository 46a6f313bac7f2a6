"""Fused optimizer step of the reference recipe.

The reference trains with `optim.Adam(lr=2e-5, weight_decay=1e-6, amsgrad=True)` and `StepLR(16000, 0.5)` stepped
once per batch (main.py:133-142, train.py:220-223).  `FusedAdam` is that optimizer as ONE hand-written kernel
(`b200_adam_step`, csrc/optimizer.cuh) over flat fp32 buffers - parameters, gradients, exp_avg, exp_avg_sq and
max_exp_avg_sq share offsets - instead of torch's nine multi_tensor_apply launches per step.

It is a `torch.optim.Optimizer`: `param_groups`, `zero_grad`, `state_dict` / `load_state_dict` (the layout of
`torch.optim.Adam`, so `TrainingState.optimizer_state` of a reference checkpoint, train.py:315, loads) and torch LR
schedulers all work.  `lr` and the step count live in device memory, so `step()` can be captured in a CUDA graph
(graphs.GraphedTrainStep); `lr_step_size > 0` applies the StepLR schedule inside the kernel, otherwise a host
scheduler may rewrite `param_groups[0]['lr']` and `sync_lr()` (called by `step()`, and by the graphed step before
each replay) forwards the new value.

`bind(net)` additionally makes the engine write weight gradients straight into the optimizer's flat gradient
buffer (and, under `parallel.DistributedUNet`, all-reduce them there), so that `step()` copies nothing.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib, ops
from .engine import GradStore


class _FlatStore(GradStore):
    """GradStore whose tensors are views of the optimizer's persistent flat gradient buffer."""

    def __init__(self, flat, slots):
        super().__init__()
        self.flat, self.slots = flat, slots

    def new(self, name, like):
        slot = self.slots.get(name)
        if slot is None:
            return super().new(name, like)
        # a NEW view object every backward: autograd adopts a gradient as `.grad` without a copy only when nothing
        # else references the tensor object
        t = self.flat[slot[0]:slot[0] + slot[1]].view(like.shape)
        self.grads[name] = t
        return t


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False,
                 lr_step_size=0, lr_gamma=1.0, model=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise RuntimeError("FusedAdam: one parameter group (the reference uses one, main.py:133-138)")
        self.lr_step_size, self.lr_gamma = int(lr_step_size), float(lr_gamma)
        self._layout = None          # built at the first step(), when the set of parameters with gradients is known
        self._net = None
        self._lr_host = None
        if model is not None:
            self.bind(model)

    # ------------------------------------------------------------------------------------------
    def bind(self, net):
        """net: the UNet or its DistributedUNet wrapper.  After the first step the engine writes gradients directly
        into this optimizer's flat buffer."""
        self._net = net
        return self

    def _module(self):
        net = self._net
        return None if net is None else (net.module if hasattr(net, "module") else net)

    def _build(self, active):
        """Flat layout over the parameters that have gradients, in the order the engine produces them (the order
        parallel.BucketedAllReduce walks, so its buckets are contiguous ranges of the same buffer)."""
        group = self.param_groups[0]
        dev = active[0].device
        module = self._module()
        names = {}
        if module is not None:
            names = {id(p): n for n, p in module.named_parameters()}
            order = module._engine().grad_order or []
            rank = {n: i for i, n in enumerate(order)}
            active = sorted(active, key=lambda p: rank.get(names.get(id(p)), len(rank)))
        offs, total = [], 0
        for p in active:
            offs.append(total)
            total += (p.numel() + 3) // 4 * 4          # every view 16-byte aligned (same rule as parallel.py)
        flat = lambda: torch.zeros(total, dtype=torch.float32, device=dev)      # noqa: E731
        L = dict(params=active, offs=offs, total=total, P=flat(), G=flat(), M=flat(), V=flat(),
                 X=flat() if group["amsgrad"] else None,
                 step=torch.zeros(1, dtype=torch.float32, device=dev), lr=torch.zeros(1, dtype=torch.float32, device=dev),
                 ticket=torch.zeros(1, dtype=torch.int32, device=dev), ids=[id(p) for p in active], gviews={}, names=names)
        old = self._layout
        with torch.no_grad():
            for p, o in zip(active, offs):
                if p.dtype != torch.float32 or not p.is_cuda:
                    raise RuntimeError("FusedAdam: fp32 CUDA parameters only")
                n = p.numel()
                view = L["P"][o:o + n].view_as(p)
                view.copy_(p)
                p.data = view                            # the parameter now lives in the flat buffer
                st = self.state[p]
                for key, buf in (("exp_avg", "M"), ("exp_avg_sq", "V"), ("max_exp_avg_sq", "X")):
                    if L[buf] is None:
                        continue
                    v = L[buf][o:o + n].view_as(p)
                    if key in st:
                        v.copy_(st[key])                 # state loaded from a checkpoint, or a previous layout
                    st[key] = v
                if "step" in st:
                    L["step"].fill_(float(st["step"]))
                st["step"] = L["step"][0]
                L["gviews"][names.get(id(p), id(p))] = (o, n)
        if old is not None:
            L["step"].copy_(old["step"])
        # one segment: the active parameters are contiguous by construction
        L["seg_begin"] = (C.c_longlong * 1)(0)
        L["seg_len"] = (C.c_longlong * 1)(total)
        self._layout = L
        self._lr_host = None
        if module is not None:
            module.invalidate_packed_weights()
            net = self._net
            if hasattr(net, "flat_provider"):            # DistributedUNet: bucketed all-reduce over OUR gradient buffer
                net.flat_provider = self._flat_for_backward
            else:
                module._grad_store_factory = self._store_for_backward

    def _accumulating(self):
        return any(p.grad is not None for p in self._layout["params"])

    def _store_for_backward(self):
        # a second backward before step() (gradient accumulation) must not overwrite the first one's gradients
        return None if self._accumulating() else _FlatStore(self._layout["G"], self._layout["gviews"])

    def _flat_for_backward(self):
        return None if self._accumulating() else self._layout["G"]

    def sync_lr(self):
        lr = float(self.param_groups[0]["lr"])
        if self._layout is not None and lr != self._lr_host:
            self._layout["lr"].fill_(lr)
            self._lr_host = lr

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        group = self.param_groups[0]
        active = [p for p in group["params"] if p.grad is not None]
        if not active:
            return loss
        capturing = torch.cuda.is_current_stream_capturing()
        L = self._layout
        stale = L is None or sorted(L["ids"]) != sorted(id(p) for p in active)
        if not stale:       # a later .to()/.cuda()/.float() of the model re-allocates the parameters
            base = L["P"].data_ptr()
            stale = any(p.data_ptr() != base + 4 * o for p, o in zip(L["params"], L["offs"]))
        if stale:
            if capturing:
                raise RuntimeError("FusedAdam: run one eager step before capturing a CUDA graph")
            self._build(active)
        L = self._layout
        for p, o in zip(L["params"], L["offs"]):           # gradients not already in the flat buffer (first step, or
            n = p.numel()                                  # an optimizer that is not bound to the model)
            if p.grad.data_ptr() != L["G"].data_ptr() + 4 * o:
                L["G"][o:o + n].view_as(p).copy_(p.grad)
        if not capturing:
            self.sync_lr()
        b1, b2 = group["betas"]
        with torch.cuda.device(L["P"].device):
            _lib.check(_lib.lib().b200_adam_step(
                L["P"].data_ptr(), L["G"].data_ptr(), L["M"].data_ptr(), L["V"].data_ptr(),
                None if L["X"] is None else L["X"].data_ptr(), L["seg_begin"], L["seg_len"], 1, L["lr"].data_ptr(),
                L["step"].data_ptr(), L["ticket"].data_ptr(), float(b1), float(b2), float(group["eps"]),
                float(group["weight_decay"]), self.lr_step_size, self.lr_gamma,
                torch.cuda.current_stream().cuda_stream), "b200_adam_step")
        ops._count(1)
        ops._ledger("adam_step_kernel", "hbm", L["total"] * 4 * (9 if L["X"] is not None else 7))
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that the engine
        # re-packs its bf16 weight images (and anything else that caches by version notices)
        for p in L["params"]:
            torch.autograd.graph.increment_version(p)
        return loss

    # ------------------------------------------------------------------------------------------
    def load_state_dict(self, state_dict):
        """Accepts the state dict of torch.optim.Adam (or of this class); values are copied into the flat buffers
        when the layout is (re)built at the next step."""
        super().load_state_dict(state_dict)
        # torch hands over the caller's tensors when dtype and device already match: own copies, so that the source
        # optimizer stepping on does not change what gets copied into the flat buffers at our next step
        for st in self.state.values():
            for k, v in list(st.items()):
                if torch.is_tensor(v):
                    st[k] = v.detach().clone()
        self._layout = None
