"""Kernel sequencing for the residual 3D U-Net: forward and backward of model.py:407-433 as a
fixed list of launches into libbrats_b200.so.

The engine owns, per input shape, a set of named zero-halo activation buffers (allocated
once, reused every step; halos stay zero because no kernel writes them) and the packed bf16
weight images (re-packed only when a parameter's version counter changes).  All launches go
to `torch.cuda.current_stream()`; nothing here synchronises the device.

Algebraic restructurings relative to the reference (all exact in real arithmetic):
  * `Conv1x1(trilinear_x2(x))` (model.py:399-402, 421) runs as `trilinear_x2(Conv1x1(x))`: the
    1x1x1 conv commutes with the interpolation, and runs on 8x fewer voxels.
  * `Conv1x1(cat([skip, up]))` (model.py:424-425) runs as one GEMM whose K loop walks the two
    source tensors; the concatenated tensor is never materialised.
  * `Conv3d(k=2, s=2)` (model.py:360-363) runs as space-to-depth followed by a 1x1x1 GEMM.
"""
from __future__ import annotations

import os

import torch

from . import ops

LEAK = 0.01


class _Plan:
    """Buffers for one (N, D, H, W) input shape."""

    def __init__(self, N, D, H, W, depth, device):
        self.N, self.device = N, device
        self.dims = [(D >> l, H >> l, W >> l) for l in range(depth)]
        self.acts = {}
        self.misc = {}
        self.generation = 0
        self.pinned = False

    def act(self, name, level, Cc):
        key = (name, level, Cc)
        t = self.acts.get(key)
        if t is None:
            D, H, W = self.dims[level]
            t = ops.act_zeros(self.N, D, H, W, Cc, self.device)
            self.acts[key] = t
        return t

    def alias(self, name, level, view):
        """Register `view` (Act.chunks of a wider buffer) under the name of a tensor of its own."""
        self.acts[(name, level, view.C)] = view
        return view

    def f32(self, name, n):
        t = self.misc.get(name)
        if t is None or t.numel() < n:
            t = torch.empty(n, dtype=torch.float32, device=self.device)
            self.misc[name] = t
        return t

    def nbytes(self):
        return sum(a.nbytes() for a in self.acts.values()) + sum(
            t.numel() * t.element_size() for t in self.misc.values() if torch.is_tensor(t))


class GradStore:
    """Where backward puts parameter gradients.  Default: one fresh fp32 tensor per parameter."""

    def __init__(self):
        self.grads = {}

    def new(self, name, like):
        t = torch.empty(like.shape, dtype=torch.float32, device=like.device)
        self.grads[name] = t
        return t

    def mark(self):
        """Called by Engine.backward at level boundaries: everything allocated so far is final."""

    def finish(self):
        """Called once at the end of backward."""


class Engine:
    def __init__(self, module):
        self.m = module
        self.depth = module.depth
        self.ch = list(module.number_of_channels)
        self.enc = list(module.encoder_layers)
        self.dec = list(module.decoder_layers)
        self.n_out = module.number_of_outputs
        if any(c % 16 or c > 128 for c in self.ch):
            raise RuntimeError("brats2019_b200: channel counts must be multiples of 16 and <= 128, got %s" % self.ch)
        if self.n_out > 4:
            raise RuntimeError("brats2019_b200: at most 4 output channels supported")
        supported = (16, 32, 64, 128)
        if any(c not in supported for c in self.ch):
            raise RuntimeError("brats2019_b200: the conv kernels exist for %s channels only, got %s" % (supported, self.ch))
        self.plans = {}
        self.max_plans = 4        # input shapes whose buffers stay resident (least recently used are freed)
        self.packed = {}
        self._packed_dirty = False
        self._checked_devices = set()
        self.grad_order = None    # parameter names in the order Engine.backward produces their gradients
        self.pack_table = ops.PackTable()
        # weight gradients run on a side stream, concurrently with the memory-bound kernels of the
        # data-gradient chain (GroupNorm backward etc.): a wgrad GEMM at 16/32 channels is bound by MMA issue
        # latency and leaves HBM and most issue slots idle, and it is off the critical path of backward.
        self.overlap_wgrad = os.environ.get("B200_NO_WGRAD_OVERLAP", "0") in ("", "0")
        # GroupNorm-backward sums folded into the epilogue of the conv that produces the gradient (no reduce pass).
        # OPT-IN (B200_GN_FOLD=1): parity-green (tests/test_layer_parity_gpu.py runs it), but measured a wash on the
        # training step (5.80 vs 5.79 ms, profiles/r02_ab_gn_fold.txt): the ~150 instructions per voxel it adds to the
        # band conv's epilogue (+50 us per launch) and the per-channel sums it moves into the apply kernel cost what
        # the dropped reduce pass saved - the level-0 backward is bound by total issue/SM occupancy, not by one kernel.
        self.fold_gn_bwd = os.environ.get("B200_GN_FOLD", "0") not in ("", "0")
        # GroupNorm statistics finished by the last CTA of the producing conv (b200_conv_run_gn) instead of a finalize
        # launch.  OPT-IN (B200_GN_FUSED_STATS=1): bit-identical, 25 launches fewer, and 0.05-0.1 ms SLOWER per forward
        # (profiles/r02_ab_last_cta_finalize.txt): every CTA pays a fence + ticket at its tail and the one CTA that
        # finishes the sums takes longer than the 4 us kernel it replaces - launches inside a CUDA graph are cheap.
        self.fused_gn_stats = os.environ.get("B200_GN_FUSED_STATS", "0") not in ("", "0")
        # depth-to-space fused into the epilogue of the k2 s2 data-gradient GEMM (EPI_D2S); 0 = separate d2s pass (A/B)
        self.fused_d2s = os.environ.get("B200_D2S_FUSED", "1") not in ("", "0")
        self._side_stream = None
        self._side_busy = False
        self._readers = {}        # id(buffer tensor) -> event recorded after the last side-stream read of it
        self._deferred = []       # side-stream weight-gradient jobs held back until the chain's next conv is enqueued
        self.fwd_pack_mode = int(os.environ.get("B200_FWD_PACK", "1") or 1)   # A/B: 0 pack -> convert on one stream, 2 swapped
        self.sched = os.environ.get("B200_BWD_SCHED", "1") not in ("", "0")   # A/B: 0 = round-2a launch order
        self.last = None          # (plan, generation) of the most recent training forward
        self.tap = None           # tests only: callable(name, Act | tensor) invoked after every backward stage, while
                                  # the (reused) gradient buffer still holds that stage's result

    # ---------------------------------------------------------------------------------
    def plan_for(self, x):
        N, Cc, D, H, W = x.shape
        q = 1 << (self.depth - 1)
        if Cc != 4 or D % q or H % q or W % q:
            raise RuntimeError("brats2019_b200: input must be (N,4,D,H,W) with D,H,W divisible by %d, got %s"
                               % (q, tuple(x.shape)))
        key = (N, D, H, W, x.device.index)
        p = self.plans.pop(key, None)
        if p is None:
            if x.device.index not in self._checked_devices:
                ops.check(ops._lib.lib().b200_device_check(x.device.index), "b200_device_check")
                self._checked_devices.add(x.device.index)
            p = _Plan(N, D, H, W, self.depth, x.device)
            # test.py crops every case to its own bounding box: a validation sweep sees many shapes, each worth GBs of
            # buffers.  Keep the most recently used few; a plan with a pending backward is never the oldest.
            while len(self.plans) >= self.max_plans and not torch.cuda.is_current_stream_capturing():
                victim = next((k for k, q in self.plans.items() if not q.pinned), None)
                if victim is None:
                    break
                self._join_side()
                self.plans.pop(victim)
        if torch.cuda.is_current_stream_capturing():
            p.pinned = True            # a CUDA graph holds raw pointers into these buffers
        self.plans[key] = p            # most recently used last
        return p

    def free_plans(self):
        """Drop every cached activation buffer (they are re-allocated on the next forward)."""
        self._join_side()
        self.plans.clear()
        self.last = None
        ops.free_workspaces()

    def invalidate_packed(self):
        """The packed bf16 weight images are refreshed when a parameter's version counter or storage changes.
        Anything that rewrites weights WITHOUT bumping the version - `p.data` edits (weight_init.py:22-27 does
        `init.kaiming_normal_(m.weight.data)`), `dist.broadcast(p.data)`, a CUDA-graph replay that contains an
        optimizer step - must call this; the model's own hooks do so for load_state_dict / apply / _apply."""
        self._packed_dirty = True

    def _params(self):
        return dict(self.m.named_parameters())

    def _pack(self, desc, kind, name, w, ci_off=0, K_real=None, N_real=None):
        """Packed bf16 operand image of one weight.  Images are registered in `self.pack_table` the first
        time they are needed and from then on refreshed all together by ONE kernel at the start of each
        forward (`_refresh_packed`); a weight that changed since then is re-packed on the spot."""
        key = (name, kind, ci_off, desc.mode, desc.epi, desc.N, desc.D, desc.H, desc.W, desc.Cin_a, desc.Cin_b, desc.Cout)
        ent = self.packed.get(key)
        ver = (w._version, w.data_ptr())
        if ent is None:
            idx, buf = self.pack_table.add(desc, kind, w, ci_off, K_real, N_real)
            ent = [ver, buf, idx]
            self.packed[key] = ent
        elif ent[0] != ver and not torch.cuda.is_current_stream_capturing():
            job = self.pack_table.jobs[ent[2]]
            if job[2].data_ptr() != w.data_ptr():
                self.pack_table.table = None          # the parameter's storage moved: rebuild the job table
            job[2] = w
            self.pack_table.repack_one(ent[2])
            ent[0] = ver
        return ent[1]

    def _refresh_packed(self):
        """One launch re-packs every registered weight image if any parameter changed (always while a CUDA
        graph is being captured: the weights differ between replays).  Returns True when the packing kernel ran."""
        if not self.packed:
            return False
        capturing = torch.cuda.is_current_stream_capturing()
        stale = capturing or self._packed_dirty
        if not stale:
            for ent in self.packed.values():
                w = self.pack_table.jobs[ent[2]][2]
                if ent[0] != (w._version, w.data_ptr()):
                    stale = True
                    break
        if stale:
            self.pack_table.run()
            for ent in self.packed.values():
                w = self.pack_table.jobs[ent[2]][2]
                ent[0] = (w._version, w.data_ptr())
            if not capturing:
                self._packed_dirty = False
        return stale

    # ---------------------------------------------------------------------------------
    # building blocks
    # ---------------------------------------------------------------------------------
    def _conv3_gn(self, P, lvl, wname, w, src, cname, Cin_real=None, gamma=None, beta=None, lrelu=True):
        """c = conv3(src) with fused GroupNorm statistics; returns (c, mean, rstd).  With gamma / beta (training) the
        finalize kernel also writes the coefficient table the backward fold reads (ops.gn_finalize)."""
        D, H, W = P.dims[lvl]
        Cout = w.shape[0]
        Cin = src.C
        desc = ops.conv_desc(ops.MODE_K3, P.N, D, H, W, Cin, Cout)
        pk = self._pack(desc, ops.W_FWD, wname, w, K_real=w.shape[1], N_real=Cout)
        c = P.act(cname, lvl, Cout)
        ctas = ops.conv_ctas(desc)
        stats = P.f32("stats:" + cname, ctas * P.N * 16)
        mean = P.f32("mean:" + cname, P.N * 8)
        rstd = P.f32("rstd:" + cname, P.N * 8)
        coef = P.f32("gnc:" + cname, P.N * 3 * Cout) if (gamma is not None and self.fold_gn_bwd) else None
        if coef is None and self.fused_gn_stats:
            # the conv's last CTA reduces the partial sums to mean / rstd: no finalize launch
            ticket = P.misc.get("gn_ticket")
            if ticket is None:
                ticket = P.misc["gn_ticket"] = torch.zeros(16, dtype=torch.int32, device=P.device)
            ops.conv_run_gn(desc, src, pk, c, stats, mean, rstd, ticket)
            return c, mean, rstd
        ops.conv_run(desc, src, pk, c, stats=stats)
        ops.gn_finalize(stats, ctas, P.N, Cout, D, H, W, mean, rstd, gamma=gamma, beta=beta, coef=coef, lrelu=lrelu)
        return c, mean, rstd

    def _residual_fwd(self, P, lvl, prefix, x_in, prm, training=False):
        """model.py:99-117 (after the optional downsample): x + lrelu(gn2(conv2(lrelu(gn1(conv1(x))))))."""
        Cc = x_in.C
        aff = (lambda n: dict(gamma=prm[prefix + n + ".weight"], beta=prm[prefix + n + ".bias"])) if training else (lambda n: {})
        c1, m1, r1 = self._conv3_gn(P, lvl, prefix + "conv1.conv1.weight", prm[prefix + "conv1.conv1.weight"], x_in,
                                    prefix + "c1", **aff("norm1"))
        a1 = ops.gn_apply(c1, m1, r1, prm[prefix + "norm1.weight"], prm[prefix + "norm1.bias"],
                          P.act(prefix + "a1", lvl, Cc), lrelu=True)
        c2, m2, r2 = self._conv3_gn(P, lvl, prefix + "conv2.conv1.weight", prm[prefix + "conv2.conv1.weight"], a1,
                                    prefix + "c2", **aff("norm2"))
        out = ops.gn_apply(c2, m2, r2, prm[prefix + "norm2.weight"], prm[prefix + "norm2.bias"],
                           P.act(prefix + "out", lvl, Cc), residual=x_in, lrelu=True)
        return out

    def _conv1(self, P, lvl, wname, w, kind, src_a, out, src_b=None, ci_off=0, residual=None):
        D, H, W = P.dims[lvl]
        Cout = out.C
        desc = ops.conv_desc(ops.MODE_K1, P.N, D, H, W, src_a.C, Cout,
                             Cin_b=0 if src_b is None else src_b.C)
        pk = self._pack(desc, kind, wname, w, ci_off=ci_off,
                        K_real=src_a.C + (0 if src_b is None else src_b.C), N_real=Cout)
        return ops.conv_run(desc, src_a, pk, out, src_b=src_b, residual=residual)

    # ---------------------------------------------------------------------------------
    # forward
    # ---------------------------------------------------------------------------------
    def forward(self, x, want_logits=False, training=False):
        P = self.plan_for(x)
        P.generation += 1
        x = x.contiguous()
        if x.dtype not in (torch.float32, torch.bfloat16):     # bf16 images are converted by the packing kernel itself
            x = x.float()
        # The weight re-pack (41 us for this model, first kernel of every training step) and the input layout conversion
        # (27 us) are independent: when the pack runs, the conversion goes to the side stream beside it, ordered behind
        # everything the main stream held BEFORE the pack was enqueued (the last readers of x16, the producer of x).
        fwd_start = torch.cuda.Event()
        fwd_start.record(torch.cuda.current_stream())
        pack_done = None
        if self.fwd_pack_mode == 2 and self.overlap_wgrad:
            # A/B (B200_FWD_PACK=2): the weight re-pack on the side stream, the input conversion on the main one
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=P.device)
            self._side_stream.wait_event(fwd_start)
            with torch.cuda.stream(self._side_stream):
                if self._refresh_packed():
                    pack_done = torch.cuda.Event()
                    pack_done.record(self._side_stream)
            repacked = False
        else:
            repacked = self._refresh_packed() and self.overlap_wgrad and self.fwd_pack_mode == 1
        prm = self._params()
        ch = self.ch
        # cat([skip, up]) of model.py:424 is ONE 2C-channel buffer per level: the encoder writes the skip
        # tensor into its first half and the up-sampling its second half (chunk-planar layout: a half is a
        # tensor of its own), so the cat conv, its data gradient and its weight gradient are single GEMMs.
        for i in range(self.depth - 1):
            if ("dec%d.cat" % i, i, 2 * ch[i]) not in P.acts:
                cat = P.act("dec%d.cat" % i, i, 2 * ch[i])
                nc = ch[i] // 8
                P.alias(self._skip_name(i), i, cat.chunks(0, nc))
                P.alias("dec%d.up" % i, i, cat.chunks(nc, 2 * nc))
        if repacked:
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream(device=P.device)
            main, side = torch.cuda.current_stream(), self._side_stream
            side.wait_event(fwd_start)
            with torch.cuda.stream(side):
                x16 = ops.pack_input(x, 16, out=P.act("x16", 0, 16))
                done = torch.cuda.Event()
                done.record(side)
            main.wait_event(done)        # (x stays alive: the main stream, which owns it, is ordered behind its last use)
        else:
            x16 = ops.pack_input(x, 16, out=P.act("x16", 0, 16))
        if pack_done is not None:
            torch.cuda.current_stream().wait_event(pack_done)
        c, m, r = self._conv3_gn(P, 0, "conv_input.weight", prm["conv_input.weight"], x16, "in.c", lrelu=False,
                                 **(dict(gamma=prm["norm_input.weight"], beta=prm["norm_input.bias"]) if training else {}))
        h = ops.gn_apply(c, m, r, prm["norm_input.weight"], prm["norm_input.bias"], P.act("in.a", 0, ch[0]),
                         lrelu=False)                                                     # model.py:412-413
        for j in range(self.enc[0]):                                                      # model.py:414
            h = self._residual_fwd(P, 0, "conv_first.%d." % j, h, prm, training)
        skips = []
        for i in range(self.depth - 1):                                                   # model.py:416-418
            skips.append(h)
            D, H, W = P.dims[i + 1]
            s2d = ops.space_to_depth(h, P.act("enc%d.s2d" % i, i + 1, 8 * ch[i]))
            wname = "encoder_convs.%d.0.downsample.0.weight" % i
            h = self._conv1(P, i + 1, wname, prm[wname], ops.W_FWD_S2D, s2d, P.act("enc%d.down" % i, i + 1, ch[i + 1]))
            for j in range(self.enc[i + 1]):
                h = self._residual_fwd(P, i + 1, "encoder_convs.%d.%d." % (i, j), h, prm, training)
        for i in reversed(range(self.depth - 1)):                                         # model.py:420-426
            wname = "upsampling.%d.1.weight" % i
            ulo = self._conv1(P, i + 1, wname, prm[wname], ops.W_FWD, h, P.act("dec%d.ulo" % i, i + 1, ch[i]))
            up = ops.upsample2x(ulo, P.act("dec%d.up" % i, i, ch[i]), lrelu=True)         # model.py:421-422
            wname = "decoder_convs1x1.%d.weight" % i
            assert up.data_ptr() == P.act("dec%d.cat" % i, i, 2 * ch[i]).chunks(ch[i] // 8, ch[i] // 4).data_ptr()
            h = self._conv1(P, i, wname, prm[wname], ops.W_FWD, P.act("dec%d.cat" % i, i, 2 * ch[i]),
                            P.act("dec%d.cc" % i, i, ch[i]))
            for j in range(self.dec[i]):
                h = self._residual_fwd(P, i, "decoder_convs.%d.%d." % (i, j), h, prm, training)
        D, H, W = P.dims[0]
        desc = ops.conv_desc(ops.MODE_K3, P.N, D, H, W, ch[0], 16, epi=ops.EPI_SIGMOID)
        w = prm["conv_output.weight"]
        pk = self._pack(desc, ops.W_FWD, "conv_output.weight", w, K_real=ch[0], N_real=self.n_out)
        probs = torch.empty((P.N, self.n_out, D, H, W), dtype=torch.float32, device=x.device)
        logits = torch.empty_like(probs) if want_logits else None
        ops.conv_run(desc, h, pk, None, bias=prm["conv_output.bias"].detach(), probs=probs, logits=logits,
                     n_out_real=self.n_out)                                               # model.py:429-431
        P.misc["final_h"] = h
        if training:
            self.last = (P, P.generation, probs)
        return (probs, logits) if want_logits else probs

    def saved_state(self):
        """(plan, generation, probs) of the training forward that just ran: what its autograd node must hand back
        to `backward` (the activations live in the plan's buffers, not in the node)."""
        return self.last

    # ---------------------------------------------------------------------------------
    # backward
    # ---------------------------------------------------------------------------------
    def _wgrad(self, P, lvl, mode, dy, x, grad, kind, ci_off=0, accumulate=False, overlap=False, defer=False):
        """Weight gradient of one conv.  overlap=True: enqueue on the side stream (after everything already on
        the current stream); the caller must not overwrite `dy` before `_wait_readers(dy)`.
        defer=True (with overlap): do not enqueue yet - `_flush_deferred()` does, after the caller has put the NEXT
        tensor-core kernel of the data-gradient chain on the main stream.  A weight-gradient CTA takes a whole SM for the
        length of the kernel (persistent, ~217 KB of shared memory), so a weight gradient that becomes runnable at the
        same moment as the chain's next conv delays that conv by up to its whole duration; behind it, it runs beside the
        memory-bound kernels that follow (trilinear adjoint, depth-to-space, GroupNorm backward)."""
        D, H, W = P.dims[lvl]
        desc = ops.wgrad_desc(mode, P.N, D, H, W, dy.C, x.C)
        overlap = overlap and self.overlap_wgrad
        wsname = "wgrad_ws_side" if overlap else "wgrad_ws"
        ws = P.misc.get(wsname)
        need = ops._lib.lib().b200_wgrad_workspace_bytes(desc) // 4
        if need == 0:
            ops.check(1, "b200_wgrad_workspace_bytes")
        if ws is None or ws.numel() < need:
            if ws is not None:
                self._join_side()          # a larger workspace replaces one the side stream may still use
            ws = torch.empty(need, dtype=torch.float32, device=P.device)
            P.misc[wsname] = ws
        if not overlap:
            ops.wgrad_run(desc, dy, x, grad, kind, ci_off=ci_off, accumulate=accumulate, workspace=ws)
            return
        job = (desc, dy, x, grad, kind, ci_off, accumulate, wsname, P)
        if defer:
            self._deferred.append(job)
        else:
            self._side_launch([job])

    def _side_launch(self, jobs):
        """Enqueue weight-gradient jobs on the side stream, ordered behind everything already on the current stream."""
        if not jobs:
            return
        P = jobs[0][8]
        if self._side_stream is None:
            self._side_stream = torch.cuda.Stream(device=P.device)
        main, side = torch.cuda.current_stream(), self._side_stream
        ready = torch.cuda.Event()
        ready.record(main)
        side.wait_event(ready)
        with torch.cuda.stream(side):
            for desc, dy, x, grad, kind, ci_off, accumulate, wsname, Pj in jobs:
                ops.wgrad_run(desc, dy, x, grad, kind, ci_off=ci_off, accumulate=accumulate, workspace=Pj.misc[wsname])
            done = torch.cuda.Event()
            done.record(side)
        for job in jobs:
            self._readers[id(job[1].t)] = done
        self._side_busy = True

    def _flush_deferred(self):
        jobs, self._deferred = self._deferred, []
        self._side_launch(jobs)

    def _wait_readers(self, buf):
        """Order the current stream after the last side-stream kernel that reads `buf` (before overwriting it)."""
        ev = self._readers.pop(id(buf.t), None)
        if ev is not None:
            torch.cuda.current_stream().wait_event(ev)

    def _join_side(self):
        if self._side_busy:
            torch.cuda.current_stream().wait_stream(self._side_stream)
            self._side_busy = False
            self._readers.clear()

    def _dgrad3(self, P, lvl, wname, w, dy, out, residual=None, fold=None):
        """Data gradient of a 3x3x3 conv.  fold = name of the conv buffer (e.g. "<prefix>c1") whose GroupNorm takes `out`
        as its dy next: its backward sums are accumulated in this conv's epilogue and returned as (gpart, ctas) for
        `_gn_bwd(folded=...)`; None is returned when the kernel this conv runs on has no such epilogue."""
        D, H, W = P.dims[lvl]
        desc = ops.conv_desc(ops.MODE_K3, P.N, D, H, W, dy.C, out.C)
        pk = self._pack(desc, ops.W_DGRAD, wname, w, K_real=w.shape[0], N_real=w.shape[1])
        coef = P.misc.get("gnc:" + fold) if (fold is not None and self.fold_gn_bwd) else None
        if coef is None or not ops.conv_supports_gnbwd(desc):
            ops.conv_run(desc, dy, pk, out, residual=residual)
            return None
        ctas = ops.conv_ctas(desc)
        gpart = P.f32("gnb:" + fold, ctas * P.N * 16)
        ops.conv_run(desc, dy, pk, out, residual=residual, stats=gpart, gnb_x=P.act(fold, lvl, out.C), gnb_coef=coef)
        return gpart, ctas

    def _gn_bwd(self, P, lvl, cname, x, dy, gamma, beta, dx, grads, gname, bname, lrelu=True, folded=None):
        Cc = x.C
        D, H, W = P.dims[lvl]
        L = ops._lib.lib()
        need = max(L.b200_gn_backward_workspace_floats(P.N, Cc), L.b200_gn_backward_folded_workspace_floats(P.N, D, H, W, Cc))
        ws = P.misc.get("gn_ws")
        if ws is None or ws.numel() < need:
            ws = torch.zeros(need, dtype=torch.float32, device=P.device)    # leading ticket words start at zero
            P.misc["gn_ws"] = ws
        dg = grads.new(gname, gamma)
        db = grads.new(bname, beta)
        if folded is not None:
            ops.gn_backward_folded(x, dy, P.misc["mean:" + cname], P.misc["rstd:" + cname], gamma, beta, folded[0],
                                   folded[1], dx, dg, db, ws, lrelu=lrelu)
        else:
            ops.gn_backward(x, dy, P.misc["mean:" + cname], P.misc["rstd:" + cname], gamma, beta, dx, dg, db, ws,
                            lrelu=lrelu)
        return dx

    def _dc_buf(self, P, lvl, Cc):
        """Next buffer of a 3-deep ring for the gradients w.r.t. conv outputs: each is read by a weight-gradient
        GEMM on the side stream, so it must not be overwritten until that GEMM is done."""
        k = P.misc.get(("dc_ring", lvl), 0)
        P.misc[("dc_ring", lvl)] = (k + 1) % 3
        buf = P.act("tmp.dc%d" % k, lvl, Cc)
        self._wait_readers(buf)
        return buf

    def _tap(self, name, value):
        if self.tap is not None:
            self._join_side()
            self.tap(name, value)

    def _residual_bwd(self, P, lvl, prefix, x_in, d_out, d_in_buf, prm, grads, d_out_folded=None, fold_next=None,
                      defer_last=False):
        """Backward of _residual_fwd.  d_out: grad w.r.t. the block output.  Returns (grad w.r.t. x_in written into
        d_in_buf - it includes the identity path, model.py:115 -, folded sums for `fold_next` or None).
        d_out_folded: the (gpart, ctas) the producer of d_out left for this block's norm2 (see _dgrad3);
        fold_next: conv buffer name whose GroupNorm consumes the returned gradient as its dy.
        defer_last: hold conv1's weight gradient back (`_wgrad(defer=True)`) - the caller enqueues a conv next."""
        Cc = x_in.C
        c1 = P.act(prefix + "c1", lvl, Cc)
        a1 = P.act(prefix + "a1", lvl, Cc)
        c2 = P.act(prefix + "c2", lvl, Cc)
        t1 = P.act("tmp.g1", lvl, Cc)
        w1n, w2n = prefix + "conv1.conv1.weight", prefix + "conv2.conv1.weight"
        dc2 = self._gn_bwd(P, lvl, prefix + "c2", c2, d_out, prm[prefix + "norm2.weight"], prm[prefix + "norm2.bias"],
                           self._dc_buf(P, lvl, Cc), grads, prefix + "norm2.weight", prefix + "norm2.bias",
                           folded=d_out_folded)
        # Each weight gradient is queued on the side stream AFTER the data-gradient conv that shares its input:
        # it then runs next to the memory-bound GroupNorm-backward kernels that follow (the two tensor-core
        # kernels cannot share an SM: both need more than half of its shared memory).
        self._tap("g:" + prefix + "d_out", d_out)
        self._tap("g:" + prefix + "dc2", dc2)
        g2 = grads.new(w2n, prm[w2n])
        f1 = self._dgrad3(P, lvl, w2n, prm[w2n], dc2, t1, fold=prefix + "c1")
        da1 = t1
        self._tap("g:" + prefix + "da1", da1)
        self._wgrad(P, lvl, 0, dc2, a1, g2, ops.G_K3, overlap=True)
        dc1 = self._gn_bwd(P, lvl, prefix + "c1", c1, da1, prm[prefix + "norm1.weight"], prm[prefix + "norm1.bias"],
                           self._dc_buf(P, lvl, Cc), grads, prefix + "norm1.weight", prefix + "norm1.bias", folded=f1)
        self._tap("g:" + prefix + "dc1", dc1)
        g1 = grads.new(w1n, prm[w1n])
        fn = self._dgrad3(P, lvl, w1n, prm[w1n], dc1, d_in_buf, residual=d_out, fold=fold_next)
        dx = d_in_buf
        self._tap("g:" + prefix + "dx", dx)
        self._wgrad(P, lvl, 0, dc1, x_in, g1, ops.G_K3, overlap=True, defer=defer_last)
        return dx, fn

    def _mark(self, grads):
        """Level boundary: a store that ships gradients (BucketedAllReduce) may launch a bucket here.  The bucket's
        weight gradients are still being produced on the side stream, so the collective is issued FROM the side stream
        (after it has been ordered behind the main stream's work so far): the NCCL stream then waits for both, and the
        main stream - the critical path of backward - is never stalled (round 1 joined the side stream here, six times
        per backward, and lost part of the wgrad / GroupNorm-backward overlap)."""
        if type(grads).mark is not GradStore.mark and self._side_busy:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._side_stream.wait_event(ev)
            with torch.cuda.stream(self._side_stream):
                grads.mark()
            return
        grads.mark()

    def backward(self, gprobs, store=None, state=None):
        """gprobs: fp32 (N, n_out, D, H, W) gradient w.r.t. the returned probabilities.
        Returns {parameter name: fp32 gradient} for the live parameters.  `store` decides where
        gradients live (GradStore: fresh tensors; parallel.BucketedAllReduce: views of one flat
        buffer, all-reduced bucket by bucket while the rest of backward still runs).
        `state`: the `saved_state()` of the forward this backward belongs to (default: the latest)."""
        state = state if state is not None else self.last
        self._deferred = []          # (a backward that raised half-way must not leave jobs for this one)
        if state is None:
            raise RuntimeError("brats2019_b200: backward without a training forward")
        P, gen, probs = state
        if gen != P.generation:
            raise RuntimeError(
                "brats2019_b200: the activations of this forward were overwritten by a later forward of the same "
                "input shape (the engine keeps ONE set of activation buffers per shape).  Call backward() before the "
                "next forward of that shape - e.g. accumulate gradients as `for xb: loss(model([xb])).backward()` "
                "instead of summing losses of several forwards first - and run eval forwards under torch.no_grad() "
                "after backward.")
        if tuple(gprobs.shape) != tuple(probs.shape):
            raise RuntimeError("brats2019_b200: gradient of shape %s for an output of shape %s"
                               % (tuple(gprobs.shape), tuple(probs.shape)))
        prm = {k: v.detach() for k, v in self._params().items()}
        ch = self.ch
        for i in range(self.depth - 1):          # gradient twins of the cat buffers: allocated on the first backward only
            if ("g.cat", i, 2 * ch[i]) not in P.acts:
                gcat = P.act("g.cat", i, 2 * ch[i])
                nc = ch[i] // 8
                P.alias("g.skip", i, gcat.chunks(0, nc))
                P.alias("g.dup", i, gcat.chunks(nc, 2 * nc))
        grads = store if store is not None else GradStore()
        gprobs = gprobs.contiguous().float()
        D, H, W = P.dims[0]
        # conv_output + sigmoid (model.py:429-431)
        dlog = P.act("g.dlogit", 0, 16)
        dbias = grads.new("conv_output.bias", prm["conv_output.bias"])
        ops.sigmoid_backward(gprobs, probs, dlog, dbias,
                             workspace=P.f32("sig_ws", ops._lib.lib().b200_sigmoid_backward_workspace_floats(P.N, D, H)))
        h_last = P.misc["final_h"]
        g_out = grads.new("conv_output.weight", prm["conv_output.weight"])
        # ping-pong gradient buffers per level: "g.A"/"g.B"
        def gbuf(name, lvl, Cc):
            buf = P.act("g." + name, lvl, Cc)
            self._wait_readers(buf)           # a side-stream weight gradient may still be reading its previous contents
            return buf

        self._tap("g:dlogit", dlog)
        # the producer of every block's output gradient leaves the sums of that block's norm2 when it is a k3 data-gradient
        # conv with the folding epilogue (`folded`), else the reduce kernel runs
        first_dec = "decoder_convs.0.%d.c2" % (self.dec[0] - 1) if self.dec[0] > 0 else None
        cur = gbuf("A", 0, ch[0])
        folded = self._dgrad3(P, 0, "conv_output.weight", prm["conv_output.weight"], dlog, cur, fold=first_dec)
        self._tap("g:final_h", cur)
        cur_name = "A"
        self._wgrad(P, 0, 0, dlog, h_last, g_out, ops.G_K3, overlap=True)

        def other(n):
            return "B" if n == "A" else "A"

        dskip = [None] * (self.depth - 1)
        # ---- decoder, levels 0 .. depth-2 (reverse of the forward order) ----
        for i in range(self.depth - 1):
            for j in reversed(range(self.dec[i])):
                prefix = "decoder_convs.%d.%d." % (i, j)
                x_in = P.act("dec%d.cc" % i, i, ch[i]) if j == 0 else P.act("decoder_convs.%d.%d.out" % (i, j - 1), i, ch[i])
                nxt = other(cur_name)
                fold_next = "decoder_convs.%d.%d.c2" % (i, j - 1) if j > 0 else None
                cur, folded = self._residual_bwd(P, i, prefix, x_in, cur, gbuf(nxt, i, ch[i]), prm, grads,
                                                 d_out_folded=folded, fold_next=fold_next,
                                                 defer_last=self.sched and j == 0)
                cur_name = nxt
            folded = None
            # cat conv (model.py:424-425): cc = W[:, :C] skip + W[:, C:] up
            # Launch order (self.sched): the chain's convs go first, every weight gradient of the level boundary goes to
            # the side stream behind them, next to the memory-bound trilinear adjoint / GroupNorm backward that follow.
            wname = "decoder_convs1x1.%d.weight" % i
            w = prm[wname]
            up = P.act("dec%d.up" % i, i, ch[i])
            self._tap("g:dec%d.cc" % i, cur)
            gcatw = grads.new(wname, w)
            if not self.sched:
                self._wgrad(P, i, 1, cur, P.act("dec%d.cat" % i, i, 2 * ch[i]), gcatw, ops.G_K1)
            self._conv1(P, i, wname, w, ops.W_DGRAD, cur, gbuf("cat", i, 2 * ch[i]))       # [dskip | dup] in one GEMM
            if self.sched:
                self._flush_deferred()
                self._wgrad(P, i, 1, cur, P.act("dec%d.cat" % i, i, 2 * ch[i]), gcatw, ops.G_K1, overlap=True)
            self._tap("g:dec%d.cat" % i, gbuf("cat", i, 2 * ch[i]))
            dskip[i] = gbuf("skip", i, ch[i])
            dup = gbuf("dup", i, ch[i])
            # lrelu + trilinear adjoint (model.py:421-422)
            dulo = ops.upsample2x_backward(dup, up, gbuf("ulo", i + 1, ch[i]), lrelu=True)
            self._tap("g:dec%d.ulo" % i, dulo)
            wname = "upsampling.%d.1.weight" % i
            w = prm[wname]
            h_lo = self._level_output(P, i + 1)
            gupw = grads.new(wname, w)
            if not self.sched:
                self._wgrad(P, i + 1, 1, dulo, h_lo, gupw, ops.G_K1)
                self._mark(grads)
            cur = self._conv1(P, i + 1, wname, w, ops.W_DGRAD, dulo, gbuf("A", i + 1, ch[i + 1]))
            if self.sched:
                self._wgrad(P, i + 1, 1, dulo, h_lo, gupw, ops.G_K1, overlap=True)
                self._mark(grads)
            self._tap("g:dec%d.h_lo" % i, cur)
            cur_name = "A"
        # At this point `cur` is the gradient w.r.t. the bottleneck output (encoder level depth-1).
        # ---- encoder, levels depth-1 .. 1 ----
        for i in reversed(range(self.depth - 1)):
            lvl = i + 1
            for j in reversed(range(self.enc[lvl])):
                prefix = "encoder_convs.%d.%d." % (i, j)
                x_in = P.act("enc%d.down" % i, lvl, ch[lvl]) if j == 0 else P.act("encoder_convs.%d.%d.out" % (i, j - 1), lvl, ch[lvl])
                nxt = other(cur_name)
                fold_next = "encoder_convs.%d.%d.c2" % (i, j - 1) if j > 0 else None
                cur, folded = self._residual_bwd(P, lvl, prefix, x_in, cur, gbuf(nxt, lvl, ch[lvl]), prm, grads,
                                                 d_out_folded=folded, fold_next=fold_next,
                                                 defer_last=self.sched and j == 0)
                cur_name = nxt
            folded = None
            wname = "encoder_convs.%d.0.downsample.0.weight" % i
            w = prm[wname]
            s2d = P.act("enc%d.s2d" % i, lvl, 8 * ch[i])
            self._tap("g:enc%d.down" % i, cur)
            gdw = grads.new(wname, w)
            if not self.sched:
                self._wgrad(P, lvl, 1, cur, s2d, gdw, ops.G_S2D)
                self._mark(grads)
            if self.fused_d2s:
                # data gradient of the k2 s2 conv written straight to the fine grid, + the skip-connection gradient from
                # the decoder (depth-to-space epilogue: no coarse 8C-channel gradient tensor, no d2s pass)
                D, H, W = P.dims[lvl]
                desc = ops.conv_desc(ops.MODE_K1, P.N, D, H, W, ch[lvl], 8 * ch[i], epi=ops.EPI_D2S)
                pk = self._pack(desc, ops.W_DGRAD_S2D, wname, w, K_real=ch[lvl], N_real=8 * ch[i])
                nxt = ops.conv_run(desc, cur, pk, gbuf("A", i, ch[i]), residual=dskip[i])
            else:
                ds2d = self._conv1(P, lvl, wname, w, ops.W_DGRAD_S2D, cur, gbuf("s2d", lvl, 8 * ch[i]))
            if self.sched:
                self._flush_deferred()
                self._wgrad(P, lvl, 1, cur, s2d, gdw, ops.G_S2D, overlap=True)
                self._mark(grads)
            if self.fused_d2s:
                cur = nxt
            else:
                # back to the fine grid, adding the skip-connection gradient from the decoder
                self._tap("g:enc%d.s2d" % i, ds2d)
                cur = ops.depth_to_space(ds2d, gbuf("A", i, ch[i]), residual=dskip[i])
            self._tap("g:enc%d.skip_total" % i, cur)
            cur_name = "A"
        # ---- level 0 head ----
        for j in reversed(range(self.enc[0])):
            prefix = "conv_first.%d." % j
            x_in = P.act("in.a", 0, ch[0]) if j == 0 else P.act("conv_first.%d.out" % (j - 1), 0, ch[0])
            nxt = other(cur_name)
            fold_next = "conv_first.%d.c2" % (j - 1) if j > 0 else "in.c"
            cur, folded = self._residual_bwd(P, 0, prefix, x_in, cur, gbuf(nxt, 0, ch[0]), prm, grads,
                                             d_out_folded=folded, fold_next=fold_next)
            cur_name = nxt
        dcin = self._gn_bwd(P, 0, "in.c", P.act("in.c", 0, ch[0]), cur, prm["norm_input.weight"], prm["norm_input.bias"],
                            gbuf(other(cur_name), 0, ch[0]), grads, "norm_input.weight", "norm_input.bias", lrelu=False,
                            folded=folded if self.enc[0] > 0 else None)
        self._tap("g:in.a", cur)
        self._tap("g:in.c", dcin)
        self._wgrad(P, 0, 0, dcin, P.act("x16", 0, 16), grads.new("conv_input.weight", prm["conv_input.weight"]),
                    ops.G_K3)
        self._flush_deferred()
        self._join_side()
        grads.finish()
        self.grad_order = list(grads.grads.keys())
        return grads.grads

    # forward-activation lookups used by backward ---------------------------------------
    def _level_output(self, P, lvl):
        """Feature map at encoder/decoder level `lvl` that was fed to upsampling[lvl-1] in forward."""
        ch = self.ch
        if lvl == self.depth - 1:
            i = lvl - 1
            return P.act("encoder_convs.%d.%d.out" % (i, self.enc[lvl] - 1), lvl, ch[lvl])
        if self.dec[lvl] > 0:
            return P.act("decoder_convs.%d.%d.out" % (lvl, self.dec[lvl] - 1), lvl, ch[lvl])
        return P.act("dec%d.cc" % lvl, lvl, ch[lvl])

    def _skip_name(self, i):
        """Name of skip_connections[i] (model.py:417), the level-i encoder output."""
        if i == 0:
            return "conv_first.%d.out" % (self.enc[0] - 1) if self.enc[0] > 0 else "in.a"
        return "encoder_convs.%d.%d.out" % (i - 1, self.enc[i] - 1)

    def _skip_tensor(self, P, i):
        """skip_connections[i] of model.py:417: the level-i encoder output."""
        ch = self.ch
        if i == 0:
            if self.enc[0] > 0:
                return P.act("conv_first.%d.out" % (self.enc[0] - 1), 0, ch[0])
            return P.act("in.a", 0, ch[0])
        return P.act("encoder_convs.%d.%d.out" % (i - 1, self.enc[i] - 1), i, ch[i])
