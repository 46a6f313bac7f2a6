"""Teacher-forced layer parity of the whole CUDA ResUNet, forward AND backward, at real model shapes.

Why this test exists.  The kernels store activations and activation gradients in bf16; the reference is fp32.
Whole-model comparisons (tests/test_model_gpu.py) therefore sit at the bf16 noise floor (~1 % of max|logit|), and
even an exact bf16-storage emulation of the pipeline (oracle.unet_logits_bf16_emulated) cannot be tighter through
depth: a one-ulp difference (fp32 summation order) in 0.01 % of the first conv's outputs perturbs every
accumulator of the next conv and flips more roundings - 1 % of elements differ after two convs, 30 % after six
(profiles/r02_layer_trace.txt).  A wrong tap in one kernel could hide under that floor.

Here nothing is propagated.  Every stage of the engine (brats2019_b200/engine.py) is checked on its own: the
stage's INPUTS are taken from the engine's own buffers (forward activations stay resident; backward gradients are
captured through `Engine.tap` while their reused buffer still holds them), the oracle's restatement of that stage
(F.conv3d / closed-form GroupNorm / trilinear stencil in fp32, rounded where the kernel stores) is applied to them,
and the result is compared with the engine's OUTPUT of that stage.  The only legitimate difference left is fp32
summation order, i.e. isolated one-ulp bf16 flips:

    stored bf16 tensors   rel-L2 <= 5e-4, every element within one bf16 ulp (+1e-4 rms), <= 1 % of elements differ
    fp32 weight gradients rel-L2 <= 2e-4 (same bf16 operands, fp32 accumulation on both sides)
    GroupNorm statistics  mean to 1e-5 (abs, relative to the group's std), rstd to 1e-5 rel
    affine / bias grads   rel-L2 <= 1e-3

Reference lines: model.py:99-117 (Residual), 407-431 (UNet.forward), loss.py:105-122 (Dice), train.py:210 (backward).
"""
import pytest
import os

import torch
import torch.nn.functional as F

from oracle import resunet_oracle as O

pytestmark = pytest.mark.gpu

ULP = 2.0 ** -7          # one bf16 ulp, relative (upper bound)


def _bf(x):
    return x.to(torch.bfloat16).float()


class _Report:
    def __init__(self, tag):
        self.tag, self.rows, self.fail = tag, [], []

    def stored(self, name, got, exp32, cap_l2=5e-4, cap_neq=0.01, ulps=1):
        """`got`: engine's stored bf16 tensor (as fp32); `exp32`: oracle value BEFORE the bf16 store.  ulps = 2 for a stage
        whose expected value is itself the sum of two rounded tensors (the skip-gradient add)."""
        exp = _bf(exp32)
        d = (got - exp).abs()
        rms = exp.pow(2).mean().sqrt().item()
        l2 = (d.norm() / exp.norm().clamp_min(1e-30)).item()
        neq = (got != exp).float().mean().item()
        over = (d > ulps * ULP * exp.abs() + 1e-4 * rms).float().mean().item()
        self.rows.append("%-44s rel-L2 %.2e  differ %.5f  >1ulp %.2e" % (name, l2, neq, over))
        if not (l2 <= cap_l2 and neq <= cap_neq and over <= 1e-5):
            self.fail.append(self.rows[-1])

    def f32(self, name, got, exp, cap=2e-4):
        l2 = ((got - exp).norm() / exp.norm().clamp_min(1e-30)).item()
        self.rows.append("%-44s rel-L2 %.2e  (fp32)" % (name, l2))
        if not l2 <= cap:
            self.fail.append(self.rows[-1])

    def finish(self):
        print("\n".join(["-- %s: %d stage checks" % (self.tag, len(self.rows))] + self.rows))
        assert not self.fail, "%s: %d stage(s) off:\n%s" % (self.tag, len(self.fail), "\n".join(self.fail))


def _conv_grads(x, w_r, dy, **kw):
    """(dX, dW) of F.conv3d(x, w_r, **kw) for output gradient dy, fp32."""
    x = x.detach().requires_grad_(True)
    w = w_r.detach().requires_grad_(True)
    y = F.conv3d(x, w, **kw)
    return torch.autograd.grad(y, (x, w), dy)


def _run(shape, with_bce, seed=11):
    import brats2019_b200 as B
    from brats2019_b200 import ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    N, D, H, W = shape
    sd = O.init_params(1337)
    # non-trivial affine parameters, so that a gamma/beta mix-up cannot hide behind weight=1, bias=0
    gsd = torch.Generator().manual_seed(5)
    for k in sd:
        if k.endswith("norm1.weight") or k.endswith("norm2.weight") or k == "norm_input.weight":
            sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=gsd)
        elif "norm" in k and k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gsd)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, 4, D, H, W, generator=g).cuda()
    t = (torch.rand(N, 3, D, H, W, generator=g) > 0.7).float().cuda()
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(sd)
    m = m.cuda().train()
    eng = m._engine()
    taps = {}
    eng.tap = lambda name, v: taps.__setitem__(name, ops.act_to_ncdhw(v) if isinstance(v, ops.Act) else v.detach().clone())
    out = m([x])
    crits = [B.Dice_loss_joint()] + ([B.BCE_Loss(bg_weight=1e-2)] if with_bce else [])
    loss = sum(c(out, [t]) for c in crits) / len(crits)
    loss.backward()
    torch.cuda.synchronize()
    eng.tap = None
    P = eng.plan_for(x)
    prm = {k: v.detach() for k, v in m.named_parameters()}
    grad = {k: v.grad for k, v in m.named_parameters() if v.grad is not None}
    wr = {k: _bf(v) for k, v in prm.items() if v.dim() == 5}         # packed (bf16) weight operands
    ch = B.DEFAULT_CFG["number_of_channels"]
    enc, dec, depth = B.DEFAULT_CFG["encoder_layers"], B.DEFAULT_CFG["decoder_layers"], B.DEFAULT_CFG["depth"]

    def act(name, lvl, C):
        return ops.act_to_ncdhw(P.acts[(name, lvl, C)])

    def stat(cname):
        return P.misc["mean:" + cname].view(N, 8), P.misc["rstd:" + cname].view(N, 8)

    R = _Report("%s%s" % (shape, " dice+bce" if with_bce else ""))

    # ------------------------------------------------------------------ forward
    def check_conv_gn(cname, lvl, C, src, w_name, gname, bname, aname, lrelu, residual=None, **kw):
        c32 = F.conv3d(src, wr[w_name], padding=1, **kw)
        c = act(cname, lvl, C)
        R.stored("fwd " + cname, c, c32)
        mean, rstd = stat(cname)
        em, er = O.gn_stats_emulated(c32)
        std = 1.0 / er
        assert ((mean - em).abs() / std).max().item() <= 1e-5, ("mean " + cname, ((mean - em).abs() / std).max().item())
        assert ((rstd - er).abs() / er).max().item() <= 1e-5, ("rstd " + cname, ((rstd - er).abs() / er).max().item())
        y = O.gn_apply_emulated(c, mean, rstd, prm[gname], prm[bname], lrelu)
        a = act(aname, lvl, C)
        R.stored("fwd " + aname, a, y if residual is None else y + residual)
        return c, a

    def check_block(prefix, lvl, C, x_in):
        _, a1 = check_conv_gn(prefix + "c1", lvl, C, x_in, prefix + "conv1.conv1.weight", prefix + "norm1.weight",
                              prefix + "norm1.bias", prefix + "a1", True)
        _, o = check_conv_gn(prefix + "c2", lvl, C, a1, prefix + "conv2.conv1.weight", prefix + "norm2.weight",
                             prefix + "norm2.bias", prefix + "out", True, residual=x_in)
        return o

    x16 = act("x16", 0, 16)
    assert torch.equal(x16[:, :4], _bf(x)) and x16[:, 4:].abs().max().item() == 0
    _, h = check_conv_gn("in.c", 0, ch[0], x16[:, :4], "conv_input.weight", "norm_input.weight", "norm_input.bias", "in.a", False)
    block_in = {}                                  # prefix -> stored input of the block (engine's)
    for j in range(enc[0]):
        block_in["conv_first.%d." % j] = h
        h = check_block("conv_first.%d." % j, 0, ch[0], h)
    skips = []
    for i in range(depth - 1):
        skips.append(h)
        wn = "encoder_convs.%d.0.downsample.0.weight" % i
        dn = act("enc%d.down" % i, i + 1, ch[i + 1])
        R.stored("fwd enc%d.down" % i, dn, F.conv3d(h, wr[wn], stride=2))
        h = dn
        for j in range(enc[i + 1]):
            block_in["encoder_convs.%d.%d." % (i, j)] = h
            h = check_block("encoder_convs.%d.%d." % (i, j), i + 1, ch[i + 1], h)
    level_out = {depth - 1: h}
    for i in reversed(range(depth - 1)):
        ulo = act("dec%d.ulo" % i, i + 1, ch[i])
        R.stored("fwd dec%d.ulo" % i, ulo, F.conv3d(h, wr["upsampling.%d.1.weight" % i]))
        up = act("dec%d.up" % i, i, ch[i])
        R.stored("fwd dec%d.up" % i, up, O.upsample_lrelu_emulated(ulo))
        cc = act("dec%d.cc" % i, i, ch[i])
        R.stored("fwd dec%d.cc" % i, cc, F.conv3d(torch.cat([skips[i], up], 1), wr["decoder_convs1x1.%d.weight" % i]))
        h = cc
        for j in range(dec[i]):
            block_in["decoder_convs.%d.%d." % (i, j)] = h
            h = check_block("decoder_convs.%d.%d." % (i, j), i, ch[i], h)
        level_out[i] = h
    logits_exp = F.conv3d(h, wr["conv_output.weight"], prm["conv_output.bias"], padding=1)
    probs = out[0].detach()
    R.f32("fwd probs", probs, torch.sigmoid(logits_exp), cap=1e-5)

    # ------------------------------------------------------------------ backward
    gp = O.dice_loss_grad_closed_form(probs, t)
    if with_bce:
        numel = float(probs.numel())
        gp = (gp + -(t / (probs + 1e-6) - 1e-2 * (1 - t) / ((1 + 1e-6) - probs)) / numel) / 2
    dlog32 = gp * probs * (1 - probs)
    dlog = taps["g:dlogit"]
    R.stored("bwd dlogit", dlog[:, :3], dlog32)
    R.f32("bwd conv_output.bias", grad["conv_output.bias"], dlog32.sum((0, 2, 3, 4)), cap=1e-3)
    dx, dw = _conv_grads(h, wr["conv_output.weight"], dlog[:, :3], padding=1)
    R.stored("bwd d(final h)", taps["g:final_h"], dx)
    R.f32("bwd conv_output.weight", grad["conv_output.weight"], dw)

    def check_block_bwd(prefix, lvl, C):
        x_in = block_in[prefix]
        d_out = taps["g:" + prefix + "d_out"]
        c1, a1, c2 = act(prefix + "c1", lvl, C), act(prefix + "a1", lvl, C), act(prefix + "c2", lvl, C)
        m2, r2 = stat(prefix + "c2")
        e, dg, db = O.gn_backward_emulated(c2, m2, r2, prm[prefix + "norm2.weight"], prm[prefix + "norm2.bias"], d_out, True)
        dc2 = taps["g:" + prefix + "dc2"]
        R.stored("bwd " + prefix + "dc2", dc2, e)
        R.f32("bwd " + prefix + "norm2.weight", grad[prefix + "norm2.weight"], dg, cap=1e-3)
        R.f32("bwd " + prefix + "norm2.bias", grad[prefix + "norm2.bias"], db, cap=1e-3)
        dxa, dw2 = _conv_grads(a1, wr[prefix + "conv2.conv1.weight"], dc2, padding=1)
        da1 = taps["g:" + prefix + "da1"]
        R.stored("bwd " + prefix + "da1", da1, dxa)
        R.f32("bwd " + prefix + "conv2.conv1.weight", grad[prefix + "conv2.conv1.weight"], dw2)
        m1, r1 = stat(prefix + "c1")
        e, dg, db = O.gn_backward_emulated(c1, m1, r1, prm[prefix + "norm1.weight"], prm[prefix + "norm1.bias"], da1, True)
        dc1 = taps["g:" + prefix + "dc1"]
        R.stored("bwd " + prefix + "dc1", dc1, e)
        R.f32("bwd " + prefix + "norm1.weight", grad[prefix + "norm1.weight"], dg, cap=1e-3)
        R.f32("bwd " + prefix + "norm1.bias", grad[prefix + "norm1.bias"], db, cap=1e-3)
        dxx, dw1 = _conv_grads(x_in, wr[prefix + "conv1.conv1.weight"], dc1, padding=1)
        R.stored("bwd " + prefix + "dx", taps["g:" + prefix + "dx"], dxx + d_out)
        R.f32("bwd " + prefix + "conv1.conv1.weight", grad[prefix + "conv1.conv1.weight"], dw1)

    dskip = {}
    for i in range(depth - 1):                     # decoder, fine to coarse (the order of Engine.backward)
        for j in reversed(range(dec[i])):
            check_block_bwd("decoder_convs.%d.%d." % (i, j), i, ch[i])
        dcc = taps["g:dec%d.cc" % i]
        up = act("dec%d.up" % i, i, ch[i])
        wn = "decoder_convs1x1.%d.weight" % i
        dcat, dw = _conv_grads(torch.cat([skips[i], up], 1), wr[wn], dcc)
        R.stored("bwd dec%d.cat [dskip|dup]" % i, taps["g:dec%d.cat" % i], dcat)
        R.f32("bwd " + wn, grad[wn], dw)
        dskip[i] = taps["g:dec%d.cat" % i][:, :ch[i]]
        dup = taps["g:dec%d.cat" % i][:, ch[i]:]
        dulo = taps["g:dec%d.ulo" % i]
        R.stored("bwd dec%d.ulo" % i, dulo, O.upsample_lrelu_backward_emulated(dup, up > 0))
        wn = "upsampling.%d.1.weight" % i
        dh, dw = _conv_grads(level_out[i + 1], wr[wn], dulo)
        R.stored("bwd dec%d.h_lo" % i, taps["g:dec%d.h_lo" % i], dh)
        R.f32("bwd " + wn, grad[wn], dw)
    for i in reversed(range(depth - 1)):           # encoder, coarse to fine
        for j in reversed(range(enc[i + 1])):
            check_block_bwd("encoder_convs.%d.%d." % (i, j), i + 1, ch[i + 1])
        ddown = taps["g:enc%d.down" % i]
        wn = "encoder_convs.%d.0.downsample.0.weight" % i
        dh, dw = _conv_grads(skips[i], wr[wn], ddown, stride=2)
        R.f32("bwd " + wn, grad[wn], dw)
        # depth-to-space epilogue (default): dh + dskip is formed in fp32 and rounded once; B200_D2S_FUSED=0: dh is stored
        # in bf16 first and the sum of the two rounded tensors is rounded again
        fused_d2s = os.environ.get("B200_D2S_FUSED", "1") not in ("", "0")
        R.stored("bwd enc%d skip total" % i, taps["g:enc%d.skip_total" % i],
                 (dh if fused_d2s else _bf(dh)) + dskip[i], ulps=1 if fused_d2s else 2)
    for j in reversed(range(enc[0])):
        check_block_bwd("conv_first.%d." % j, 0, ch[0])
    mi, ri = stat("in.c")
    e, dg, db = O.gn_backward_emulated(act("in.c", 0, ch[0]), mi, ri, prm["norm_input.weight"], prm["norm_input.bias"],
                                       taps["g:in.a"], False)
    R.stored("bwd in.c", taps["g:in.c"], e)
    R.f32("bwd norm_input.weight", grad["norm_input.weight"], dg, cap=1e-3)
    R.f32("bwd norm_input.bias", grad["norm_input.bias"], db, cap=1e-3)
    _, dw = _conv_grads(x16[:, :4], wr["conv_input.weight"], taps["g:in.c"], padding=1)
    R.f32("bwd conv_input.weight", grad["conv_input.weight"], dw)
    assert len(grad) == 86 and all(n not in grad for n in m.dead_parameter_names())
    R.finish()
    del m, taps
    torch.cuda.empty_cache()


# (3,8,16,24): odd batch, a single slice at the deepest level; (1,24,40,72): ragged tiles at every level;
# (2,128,128,128): BASELINE config 3 exactly (batch 2 x 4x128^3 training step).
@pytest.mark.parametrize("shape,with_bce", [((2, 16, 24, 32), True), ((3, 8, 16, 24), False), ((1, 24, 40, 72), False),
                                            ((2, 128, 128, 128), False)])
def test_every_stage_matches_the_oracle_on_the_engines_own_inputs(shape, with_bce):
    _run(shape, with_bce)


@pytest.mark.parametrize("shape", [(2, 16, 24, 32), (1, 24, 40, 72), (2, 64, 64, 128)])
def test_every_stage_with_the_groupnorm_backward_fold(shape, monkeypatch):
    """The opt-in path (B200_GN_FOLD=1): GroupNorm-backward sums accumulated in the epilogue of the data-gradient conv
    that produces the gradient (band kernel at 16 channels, marching kernel at 32), dgamma / dbeta from sums the apply
    kernel writes.  Same per-stage bar."""
    monkeypatch.setenv("B200_GN_FOLD", "1")
    _run(shape, False)
