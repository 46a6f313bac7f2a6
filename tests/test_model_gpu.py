"""GPU parity of the CUDA ResUNet (through the nn.Module drop-in, i.e. through the C ABI)
against (a) golden vectors produced by the real reference and (b) the fp32 oracle at larger
sizes.

Three layers of evidence, tightest first:

 1. tests/test_layer_parity_gpu.py - every stage of the engine, forward and backward, against the oracle's
    restatement applied to the engine's own inputs: one-ulp-exact (rel-L2 <= 5e-4 per stored tensor, <= 2e-4 per
    weight gradient).  That is the pin that can see a wrong tap.
 2. Here, whole model vs the bf16-storage EMULATION of itself (oracle.train_step_bf16_emulated): not limited by
    the storage format, but by the chaotic amplification of one-ulp rounding flips through 60 stored tensors
    (profiles/r02_layer_trace.txt), so it can only be required to be CLOSER than the fp32 comparison:
    mean logit error <= 0.85x and (from 32^3 voxels up) median gradient error <= 0.9x of the respective error
    against fp32.
 3. Here, whole model vs the fp32 reference (golden vectors of the real reference / the fp32 oracle).  The
    kernels store activations in bf16 and accumulate in fp32; the reference is fp32.  The yardstick for
    "bf16-correct" is the reference arithmetic itself under torch.autocast(bfloat16) on the same GPU
    (even that differs from fp32 by ~1% of max|logit| and flips ~0.3% of thresholded voxels).
    Every statistic must satisfy BOTH an absolute cap and "<= 1.5x the yardstick's own error" (2x for the two
    max-over-all-voxels statistics, which are single-sample extremes):

    logits     max-abs err <= 3% of max|logit|;  mean-abs err <= 2% of mean|logit|
    probs      max-abs err <= 0.08
    masks      Dice(mask_cuda, mask_ref) >= 0.98 per channel (threshold 0.5, test.py:144)
    metrics.Dice(mask, target) (metrics.py:108-133): |cuda - ref| <= 5e-3 per sample and channel (<= 2.5e-3 at
               the config-3 shape; measured 1.6e-3 there - the north star's 1e-3 is not reachable with bf16
               storage, the yardstick misses it too)
    Dice loss  |loss - ref| <= 1e-3   (north star: 1e-3; measured <= 4e-4)
    gradients  per-tensor relative L2 err: max <= 0.20, median <= 0.06 (<= 0.08 / 0.02 at the config-3 shape)
(profiles/r02_parity_report.txt holds the measured numbers next to the yardstick.)
"""
import numpy as np
import pytest
import torch

from oracle import resunet_oracle as O

pytestmark = pytest.mark.gpu

CAP = dict(logit_max=0.03, logit_mean=0.02, prob=0.08, mask_dice=0.98, loss=1e-3, grad_max=0.20, grad_median=0.06,
           metric=5e-3)
YARD = 1.5


def _model(sd, dev="cuda"):
    import brats2019_b200 as B
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(sd)
    return m.to(dev)


def _fwd_stats(logits, ref_logits):
    ref_logits = ref_logits.to(logits.device)
    err = (logits - ref_logits).abs()
    p, pr = torch.sigmoid(logits), torch.sigmoid(ref_logits)
    ma, mb = p > 0.5, pr > 0.5
    dice = []
    for c in range(p.shape[1]):
        den = ma[:, c].sum().item() + mb[:, c].sum().item()
        dice.append(2.0 * (ma[:, c] & mb[:, c]).sum().item() / den if den else 1.0)
    return dict(logit_max=err.max().item() / ref_logits.abs().max().item(),
                logit_mean=err.mean().item() / ref_logits.abs().mean().item(),
                prob=(p - pr).abs().max().item(), mask_dice=min(dice))


def _yardstick(sd, x, t):
    """The reference arithmetic under bf16 autocast on this GPU, vs itself in fp32."""
    sdc = {k: v.cuda() for k, v in sd.items()}
    xc, tc = x.cuda(), t.cuda()
    ref_logits = O.unet_logits(sdc, xc)
    _, _, g32 = O.train_step(sdc, xc, tc)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        lb = O.unet_logits(sdc, xc).float()
        _, _, gbf = O.train_step(sdc, xc, tc)
    rel = torch.tensor([((gbf[k].float() - g32[k]).norm() / g32[k].norm().clamp_min(1e-20)).item() for k in g32])
    st = _fwd_stats(lb, ref_logits)
    st.update(grad_max=rel.max().item(), grad_median=rel.median().item())
    return st


def _grad_rels(model, ref_grads):
    out = {}
    for n, p in model.named_parameters():
        if p.grad is not None and n in ref_grads:
            ref = torch.as_tensor(ref_grads[n]).to(p.grad.device).float()
            out[n] = ((p.grad - ref).norm() / ref.norm().clamp_min(1e-20)).item()
    return out


def _emulated(sd, x, t, with_bce=False):
    """The bf16-storage emulation of the CUDA pipeline (oracle), fp32 torch ops on this GPU."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    loss, probs, logits, grads = O.train_step_bf16_emulated(sdc, x.cuda(), t.cuda(), with_bce=with_bce)
    return loss, logits, grads


def _check_closer_to_emulation(model, logits, ref_logits, ref_grads, em_logits, em_grads, tag):
    """Evidence layer 2 of the module docstring."""
    ref_logits = ref_logits.to(logits.device)
    e_ref = (logits - ref_logits).abs().mean().item()
    e_em = (logits - em_logits).abs().mean().item()
    r_ref = torch.tensor(list(_grad_rels(model, ref_grads).values())).median().item()
    r_em = torch.tensor(list(_grad_rels(model, em_grads).values())).median().item()
    print(tag, "mean |dlogit| vs fp32 %.5f vs bf16-emulated %.5f | median grad rel-L2 vs fp32 %.5f vs emulated %.5f" % (
        e_ref, e_em, r_ref, r_em))
    assert e_em <= 0.85 * e_ref, (tag, e_em, e_ref)
    # gradients: at tiny volumes (2^3 voxels at the deepest level) single roundings dominate both numbers
    big = logits[0, 0].numel() >= 32 ** 3
    assert r_em <= (0.9 if big else 1.5) * r_ref, (tag, r_em, r_ref)


def _check_metric(probs, ref_probs, target, tag, cap=None):
    """metrics.Dice (metrics.py:108-133) of the thresholded masks against the target: CUDA vs reference."""
    ref_probs, target = ref_probs.to(probs.device), target.to(probs.device)
    a = O.dice_metric(O.threshold_masks(probs), target > 0.5)
    b = O.dice_metric(O.threshold_masks(ref_probs), target > 0.5)
    d = (a - b).abs().max().item()
    print(tag, "metrics.Dice(mask, target): max |cuda - ref| %.2e (ref values %s)" % (d, [round(v, 4) for v in b.flatten().tolist()]))
    assert d <= (cap or CAP["metric"]), (tag, d)


def _check_forward(probs, logits, ref_logits, yard, tag):
    st = _fwd_stats(logits, ref_logits)
    print(tag, "forward", {k: round(v, 5) for k, v in st.items()}, "yardstick", {k: round(yard[k], 5) for k in st})
    assert torch.isfinite(logits).all() and torch.isfinite(probs).all()
    assert (probs - torch.sigmoid(logits)).abs().max().item() < 1e-5
    for k in ("logit_max", "logit_mean", "prob"):
        assert st[k] <= CAP[k], (tag, k, st[k])
        assert st[k] <= (YARD if k == "logit_mean" else 2.0) * yard[k] + 1e-3, (tag, k, st[k], yard[k])
    assert st["mask_dice"] >= CAP["mask_dice"], (tag, st["mask_dice"])
    assert 1 - st["mask_dice"] <= YARD * (1 - yard["mask_dice"]) + 2e-3, (tag, st["mask_dice"], yard["mask_dice"])


def _check_grads(model, ref_grads, yard, tag):
    rels = []
    dead = set(model.dead_parameter_names())
    for n, p in model.named_parameters():
        if n in dead:
            assert p.grad is None, "dead parameter %s received a gradient" % n
            continue
        assert p.grad is not None, "no gradient for %s" % n
        assert torch.isfinite(p.grad).all(), n
        if n not in ref_grads:
            continue
        ref = torch.as_tensor(ref_grads[n]).to(p.grad.device).float()
        rels.append((((p.grad - ref).norm() / ref.norm().clamp_min(1e-20)).item(), n))
    t = torch.tensor([r for r, _ in rels])
    mx, med = t.max().item(), t.median().item()
    print(tag, "gradient rel-L2 max %.4f median %.4f (worst %s); yardstick max %.4f median %.4f" % (
        mx, med, max(rels)[1], yard["grad_max"], yard["grad_median"]))
    assert mx <= CAP["grad_max"] and mx <= YARD * yard["grad_max"] + 1e-2, (tag, max(rels), yard["grad_max"])
    if len(rels) >= 40:
        assert med <= CAP["grad_median"] and med <= YARD * yard["grad_median"] + 5e-3, (tag, med, yard["grad_median"])


@pytest.mark.parametrize("case", ["cube16_b1", "box16x24x32_b2"])
def test_forward_and_backward_match_reference_golden(golden, case):
    import brats2019_b200 as B
    g = golden(case)
    sd = O.init_params(int(g["weight_seed"]))
    m = _model(sd)
    x = torch.from_numpy(g["x"]).cuda()
    t = torch.from_numpy(g["target"]).float().cuda()
    # forward (eval / no_grad path) with logits
    m.eval()
    (probs,), logits = m([x], return_logits=True)
    yard = _yardstick(sd, torch.from_numpy(g["x"]), torch.from_numpy(g["target"]).float())
    _check_forward(probs, logits, torch.from_numpy(g["logits"]), yard, case)
    # training path: Dice loss + backward
    m.train()
    out = m([x])
    crit = B.Dice_loss_joint(index=0, priority=1)
    loss = crit(out, [t])
    assert abs(loss.item() - float(g["dice"])) <= CAP["loss"], (loss.item(), float(g["dice"]))
    loss.backward()
    ref = {k[len("gdice::"):]: g[k] for k in g.files if k.startswith("gdice::")}
    _check_grads(m, ref, yard, case)
    _check_metric(out[0].detach(), torch.from_numpy(g["probs"]), t, case)
    _, em_logits, em_grads = _emulated(sd, torch.from_numpy(g["x"]), t)
    _check_closer_to_emulation(m, logits, torch.from_numpy(g["logits"]), ref, em_logits, em_grads, case)
    names = [str(n) for n in g["live_names"]]
    prm = dict(m.named_parameters())
    got = np.array([prm[n].grad.double().norm().item() for n in names])
    rel = np.abs(got - g["grad_norm_dice"]) / np.maximum(g["grad_norm_dice"], 1e-20)
    print(case, "grad-norm rel err: max %.4f median %.4f" % (rel.max(), np.median(rel)))
    assert rel.max() <= 0.15 and np.median(rel) <= 0.03, list(zip(names, rel))[int(rel.argmax())]
    # full criterion of the trainer: mean(Dice, BCE) (main.py:126-128, train.py:203-205)
    m.zero_grad()
    out = m([x])
    l2 = (crit(out, [t]) + B.BCE_Loss(index=0, bg_weight=1e-2)(out, [t])) / 2
    l2.backward()
    ref = {k[len("gboth::"):]: g[k] for k in g.files if k.startswith("gboth::")}
    _check_grads(m, ref, yard, case + " dice+bce")


# (3, 8, 16, 24): odd batch, one slice at the deepest level (D/8 = 1); (1, 24, 40, 72): no dimension a multiple
# of 16 at level 1+, ragged tiles everywhere; (1, 128, 128, 128): one volume of the benchmark size (config 1/3).
# (2, 128, 128, 128): BASELINE config 3 exactly.
@pytest.mark.parametrize("shape", [(1, 32, 32, 32), (2, 32, 48, 64), (3, 8, 16, 24), (1, 24, 40, 72), (1, 128, 128, 128),
                                   (2, 128, 128, 128)])
def test_matches_oracle_at_larger_sizes(shape):
    import brats2019_b200 as B
    N, D, H, W = shape
    sd = O.init_params(1337)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, 4, D, H, W, generator=g)
    t = (torch.rand(N, 3, D, H, W, generator=g) > 0.7).float()
    loss_ref, probs_ref, grads_ref = O.train_step(sd, x, t)
    logits_ref = O.unet_logits(sd, x)
    m = _model(sd)
    (probs,), logits = m([x.cuda()], return_logits=True)
    yard = _yardstick(sd, x, t)
    _check_forward(probs, logits, logits_ref, yard, "oracle %s" % (shape,))
    out = m([x.cuda()])
    loss = B.Dice_loss_joint()(out, [t.cuda()])
    assert abs(loss.item() - loss_ref.item()) <= CAP["loss"]
    loss.backward()
    _check_grads(m, grads_ref, yard, "oracle %s" % (shape,))
    big = D * H * W >= 128 ** 3
    _check_metric(out[0].detach(), probs_ref, t, "oracle %s" % (shape,), cap=2.5e-3 if big else None)
    if big:          # the benchmark shape: enough voxels per weight for much tighter gradient statistics
        rels = torch.tensor(list(_grad_rels(m, grads_ref).values()))
        assert rels.max().item() <= 0.08 and rels.median().item() <= 0.02, (rels.max().item(), rels.median().item())
    _, em_logits, em_grads = _emulated(sd, x, t)
    _check_closer_to_emulation(m, logits, logits_ref, grads_ref, em_logits, em_grads, "oracle %s" % (shape,))
    del m
    torch.cuda.empty_cache()


def test_eval_and_train_paths_agree_and_are_repeatable():
    sd = O.init_params(3)
    m = _model(sd)
    x = torch.randn(1, 4, 16, 16, 16, device="cuda")
    with torch.no_grad():
        a = m([x])[0].clone()
    b = m([x])[0].detach().clone()
    c = m([x])[0].detach().clone()
    assert torch.equal(a, b) and torch.equal(b, c)      # deterministic, same kernels both paths


def test_optimizer_steps_reduce_the_loss():
    import brats2019_b200 as B
    torch.manual_seed(0)
    m = _model(O.init_params(5))
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)      # dead parameters have grad None: Adam skips them
    x = torch.randn(2, 4, 16, 16, 16, device="cuda")
    t = (torch.rand(2, 3, 16, 16, 16, device="cuda") > 0.7).float()
    crit = B.Dice_loss_joint()
    losses = []
    for _ in range(8):
        opt.zero_grad()
        loss = crit(m([x]), [t])
        loss.backward()
        opt.step()
        losses.append(loss.item())
    print("losses", losses)
    assert losses[-1] < losses[0]


def test_rejects_bad_inputs():
    m = _model(O.init_params(1))
    with pytest.raises(RuntimeError):
        m([torch.zeros(1, 4, 12, 16, 16, device="cuda")])      # 12 not divisible by 8
    with pytest.raises(RuntimeError):
        m([torch.zeros(1, 4, 16, 16, 16)])                     # CPU tensor: no fallback


def _oracle_gpu_fp32(sd, x):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sdc = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        return O.unet_logits(sdc, x)


def test_full_brats_volume_inference_config2():
    """BASELINE config 2: one 4x240x240x155 volume zero-padded to 4x240x240x160 (test.py:93,
    loader_helper.py:99-103), eval / no_grad.  Checker: the oracle in fp32 on the same GPU."""
    sd = O.init_params(1337)
    g = torch.Generator().manual_seed(21)
    x = torch.zeros(1, 4, 240, 240, 160)
    x[..., :155] = torch.randn(1, 4, 240, 240, 155, generator=g)
    x = x.cuda()
    m = _model(sd).eval()
    (probs,), logits = m([x], return_logits=True)
    ref = _oracle_gpu_fp32(sd, x)
    st = _fwd_stats(logits, ref)
    print("config2 240x240x160", {k: round(v, 5) for k, v in st.items()})
    assert torch.isfinite(logits).all()
    assert st["logit_max"] <= CAP["logit_max"] and st["logit_mean"] <= CAP["logit_mean"]
    assert st["prob"] <= CAP["prob"] and st["mask_dice"] >= CAP["mask_dice"]
    del m, probs, logits, ref
    torch.cuda.empty_cache()


def test_reference_training_patch_shape_144x144x128():
    """main.py:111: SimpleReader patch 144x144x128, batch 1 (main.py:22 default): fwd + Dice + bwd."""
    import brats2019_b200 as B
    sd = O.init_params(1337)
    g = torch.Generator().manual_seed(22)
    x = torch.randn(1, 4, 144, 144, 128, generator=g).cuda()
    t = (torch.rand(1, 3, 144, 144, 128, generator=g) > 0.7).float().cuda()
    sdc = {k: v.cuda() for k, v in sd.items()}
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    loss_ref, _, grads_ref = O.train_step(sdc, x, t)
    m = _model(sd).train()
    loss = B.Dice_loss_joint()(m([x]), [t])
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) <= CAP["loss"]
    rels = torch.tensor([((p.grad - grads_ref[n]).norm() / grads_ref[n].norm().clamp_min(1e-20)).item()
                         for n, p in m.named_parameters() if p.grad is not None])
    print("patch 144x144x128 grad rel-L2 max %.4f median %.4f" % (rels.max().item(), rels.median().item()))
    assert rels.max().item() <= CAP["grad_max"] and rels.median().item() <= CAP["grad_median"]
    del m
    torch.cuda.empty_cache()


def test_losses_accept_contiguous_views_that_are_not_16_byte_aligned():
    """`.contiguous()` is a no-op for a sliced batch whose storage offset is not a multiple of 4 floats;
    the 16-byte load path must not be taken for it (and odd spatial sizes must work)."""
    import brats2019_b200 as B
    for S, off in ((1000, 1), (1000, 2), (999, 0), (1001, 3)):
        g = torch.Generator().manual_seed(S + off)
        p = torch.rand(off + 2 * 3 * S, generator=g).cuda()[off:].view(2, 3, S, 1, 1).requires_grad_(True)
        t = (torch.rand(off + 2 * 3 * S, generator=g) > 0.6).float().cuda()[off:].view(2, 3, S, 1, 1)
        assert p.is_contiguous() and (p.data_ptr() % 16 != 0 or S % 4)
        for crit, ref in ((B.Dice_loss_joint(), O.dice_loss_joint), (B.BCE_Loss(bg_weight=1e-2), lambda a, b: O.bce_loss(a, b, bg_weight=1e-2))):
            loss = crit([p], [t])
            want = ref([p.detach().cpu().double()], [t.cpu().double()])
            assert abs(loss.item() - want.item()) <= 1e-5 * max(1.0, abs(want.item())), (S, off, loss.item(), want.item())
            (gp,) = torch.autograd.grad(loss, p)
            pr = p.detach().cpu().double().requires_grad_(True)
            (gr,) = torch.autograd.grad(ref([pr], [t.cpu().double()]), pr)
            assert (gp.cpu().double() - gr).abs().max().item() <= 1e-4 * gr.abs().max().item() + 1e-9
