"""GPU parity of the CUDA ResUNet (through the nn.Module drop-in, i.e. through the C ABI)
against (a) golden vectors produced by the real reference and (b) the CPU oracle at larger
sizes.  Tolerances (bf16 storage / fp32 accumulate vs the reference's fp32):

    logits     max-abs error <= 3% of max|logit|,  mean-abs error <= 1% of mean|logit|... (TOL_*)
    probs      max-abs error <= 0.03
    masks      Dice(mask_cuda, mask_ref) >= 0.99 per channel (thresholded at 0.5, test.py:144)
    Dice loss  |loss - ref| <= 2e-3
    gradients  per-tensor relative L2 error <= 5e-2 (<= 8e-2 for tensors with < 64 elements)
"""
import numpy as np
import pytest
import torch

from oracle import resunet_oracle as O

pytestmark = pytest.mark.gpu

TOL_LOGIT_MAX, TOL_LOGIT_MEAN, TOL_PROB, TOL_MASK_DICE, TOL_LOSS, TOL_GRAD = 0.03, 0.01, 0.03, 0.99, 2e-3, 5e-2


def _model(sd, dev="cuda"):
    import brats2019_b200 as B
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(sd)
    return m.to(dev)


def _check_forward(probs, logits, ref_logits, tag):
    ref_logits = ref_logits.to(logits.device)
    ref_probs = torch.sigmoid(ref_logits)
    err = (logits - ref_logits).abs()
    mx, mean = ref_logits.abs().max().item(), ref_logits.abs().mean().item()
    msg = "%s: logit err max %.4f (ref max %.2f) mean %.5f (ref mean %.3f); prob err max %.4f" % (
        tag, err.max().item(), mx, err.mean().item(), mean, (probs - ref_probs).abs().max().item())
    print(msg)
    assert torch.isfinite(logits).all() and torch.isfinite(probs).all(), msg
    assert err.max().item() <= TOL_LOGIT_MAX * mx, msg
    assert err.mean().item() <= TOL_LOGIT_MEAN * mean, msg
    assert (probs - ref_probs).abs().max().item() <= TOL_PROB, msg
    ma, mb = probs > 0.5, ref_probs > 0.5
    for c in range(probs.shape[1]):
        inter = (ma[:, c] & mb[:, c]).sum().item()
        den = ma[:, c].sum().item() + mb[:, c].sum().item()
        d = 2.0 * inter / den if den else 1.0
        assert d >= TOL_MASK_DICE, "%s: mask dice channel %d = %.5f" % (tag, c, d)


def _check_grads(model, ref_grads, tag):
    worst = []
    dead = set(model.dead_parameter_names())
    for n, p in model.named_parameters():
        if n in dead:
            assert p.grad is None, "dead parameter %s received a gradient" % n
            continue
        assert p.grad is not None, "no gradient for %s" % n
        if n not in ref_grads:
            continue
        ref = torch.as_tensor(ref_grads[n]).to(p.grad.device).float()
        rel = ((p.grad - ref).norm() / ref.norm().clamp_min(1e-20)).item()
        worst.append((rel, n, ref.numel()))
    worst.sort(reverse=True)
    print(tag, "worst gradient rel-L2:", ["%s %.4f" % (n, r) for r, n, _ in worst[:6]])
    for rel, n, numel in worst:
        tol = TOL_GRAD if numel >= 64 else 8e-2
        assert rel <= tol, "%s: gradient %s rel L2 error %.4f > %.3f" % (tag, n, rel, tol)
    return worst


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
    _check_forward(probs, logits, torch.from_numpy(g["logits"]), case)
    # training path: Dice loss + backward
    m.train()
    out = m([x])
    crit = B.Dice_loss_joint(index=0, priority=1)
    loss = crit(out, [t])
    assert abs(loss.item() - float(g["dice"])) <= TOL_LOSS, (loss.item(), float(g["dice"]))
    loss.backward()
    ref = {k[len("gdice::"):]: g[k] for k in g.files if k.startswith("gdice::")}
    _check_grads(m, ref, case)
    names = [str(n) for n in g["live_names"]]
    prm = dict(m.named_parameters())
    got = np.array([prm[n].grad.double().norm().item() for n in names])
    rel = np.abs(got - g["grad_norm_dice"]) / np.maximum(g["grad_norm_dice"], 1e-20)
    print(case, "grad-norm rel err: max %.4f median %.4f" % (rel.max(), np.median(rel)))
    assert rel.max() <= 6e-2, list(zip(names, rel))[int(rel.argmax())]
    # full criterion of the trainer: mean(Dice, BCE) (main.py:126-128, train.py:203-205)
    m.zero_grad()
    out = m([x])
    l2 = (crit(out, [t]) + B.BCE_Loss(index=0, bg_weight=1e-2)(out, [t])) / 2
    l2.backward()
    ref = {k[len("gboth::"):]: g[k] for k in g.files if k.startswith("gboth::")}
    _check_grads(m, ref, case + " dice+bce")


@pytest.mark.parametrize("shape", [(1, 32, 32, 32), (2, 32, 48, 64)])
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
    _check_forward(probs, logits, logits_ref, "oracle %s" % (shape,))
    out = m([x.cuda()])
    loss = B.Dice_loss_joint()(out, [t.cuda()])
    assert abs(loss.item() - loss_ref.item()) <= TOL_LOSS
    loss.backward()
    _check_grads(m, grads_ref, "oracle %s" % (shape,))


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
