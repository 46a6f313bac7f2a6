"""CPU: the oracle restatement reproduces the committed outputs of the real reference."""
import hashlib

import numpy as np
import pytest
import torch

from oracle import resunet_oracle as O

CASES = ["cube16_b1", "box16x24x32_b2"]


def _digest(sd):
    m = hashlib.sha256()
    for k, v in sd.items():
        m.update(k.encode())
        m.update(v.numpy().tobytes())
    return m.hexdigest()


def test_param_contract():
    shapes = O.param_shapes()
    assert len(shapes) == 93                                  # 93 tensors
    assert sum(int(np.prod(s)) for _, s in shapes) == 5427955  # 5,427,955 parameters
    dead = set(O.dead_param_names())
    assert len(dead) == 7
    live = [(n, s) for n, s in shapes if n not in dead]
    assert len(live) == 86 and sum(int(np.prod(s)) for _, s in live) == 4509939


@pytest.mark.parametrize("case", CASES)
def test_oracle_matches_reference_golden(golden, case):
    g = golden(case)
    sd = O.init_params(int(g["weight_seed"]))
    assert _digest(sd) == str(g["params_sha256"]), "torch CPU RNG stream differs from the golden's"
    x = torch.from_numpy(g["x"])
    t = torch.from_numpy(g["target"]).float()
    logits = O.unet_logits(sd, x)
    np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=2e-5)
    probs = torch.sigmoid(logits)
    np.testing.assert_allclose(probs.numpy(), g["probs"], rtol=0, atol=1e-6)
    assert abs(O.dice_loss_joint([probs], [t]).item() - float(g["dice"])) < 1e-6
    assert abs(O.bce_loss([probs], [t], bg_weight=1e-2).item() - float(g["bce"])) < 1e-6

    names = [str(n) for n in g["live_names"]]
    for with_bce, key, nk in ((False, "gdice::", "grad_norm_dice"), (True, "gboth::", "grad_norm_both")):
        loss, _, grads = O.train_step(sd, x, t, with_bce=with_bce)
        assert list(grads.keys()) == names
        norms = np.array([grads[k].double().norm().item() for k in names])
        np.testing.assert_allclose(norms, g[nk], rtol=2e-4, atol=1e-9)
        for k in g.files:
            if k.startswith(key):
                ref = g[k]
                got = grads[k[len(key):]].numpy()
                assert np.abs(got - ref).max() <= 2e-4 * max(np.abs(ref).max(), 1e-12) + 1e-9, k


def test_trilinear_stencil_is_interpolate():
    x = torch.randn(2, 3, 4, 6, 5)
    np.testing.assert_allclose(O.trilinear_x2_explicit(x).numpy(), O.trilinear_x2(x).numpy(), atol=1e-6)


def test_dice_closed_form_gradient():
    torch.manual_seed(0)
    p = torch.rand(2, 3, 4, 4, 4, requires_grad=True)
    t = (torch.rand(2, 3, 4, 4, 4) > 0.6).float()
    O.dice_loss_joint([p], [t]).backward()
    np.testing.assert_allclose(O.dice_loss_grad_closed_form(p.detach(), t).numpy(), p.grad.numpy(), rtol=1e-4, atol=1e-8)


def test_tiling_restatement_matches_reference_loader_helper(golden):
    """oracle.tile_* / predict_tiled vs outputs of the reference's loader_helper (tests/golden/tiling.npz)."""
    g = golden("tiling")
    data = torch.from_numpy(g["data"])
    tile, center, border = tuple(int(v) for v in g["tile"]), tuple(int(v) for v in g["center"]), tuple(int(v) for v in g["border"])
    seen = []
    state = {"n": 0}
    grid = [int(np.ceil(j / i)) for i, j in zip(center, data.shape[2:])]

    def fn(t):
        n = state["n"]
        state["n"] += 1
        seen.append(t.numpy().copy())
        i, rem = divmod(n, grid[1] * grid[2])
        j, k = divmod(rem, grid[2])
        return t * 2.0 + (i * 100 + j * 10 + k)

    out = O.predict_tiled(fn, data, data.shape, tile, center, border)
    np.testing.assert_array_equal(np.stack(seen), g["tiles"])
    np.testing.assert_array_equal(out.numpy(), g["out"])


def test_product_tiler_matches_oracle_tiler_on_cpu_tensors():
    """brats2019_b200.inference.predict_tiled cuts the same windows and writes the same centres as the
    reference procedure (checked with a stand-in model on CPU tensors: only tensor indexing runs)."""
    from brats2019_b200 import inference as I
    torch.manual_seed(0)
    x = torch.randn(1, 4, 21, 13, 10)
    tile, center, border = (10, 8, 6), (4, 4, 2), (3, 2, 2)

    class Fake:
        def __call__(self, lst):
            t = lst[0]
            return [torch.stack([t[:, 0] * 2 + t[:, 1], t[:, 2] - t[:, 3], t[:, 0] * t[:, 3]], dim=1)]

    ref = O.predict_tiled(lambda t: Fake()([t])[0], x, (1, 3, 21, 13, 10), tile, center, border)
    for bt in (1, 3):
        got = I.predict_tiled(Fake(), x, 3, tile, center, border, batch_tiles=bt)
        np.testing.assert_array_equal(got.numpy(), ref.numpy())
    # padding rule of test.py:92-99
    xp, l, r = I.pad_to_multiple(x, 16)
    assert xp.shape[2:] == (32, 16, 16) and torch.equal(I.unpad(xp, l, r), x)
    assert [I.closest_to_k(v, 16) for v in (155, 240, 160, 16)] == [160, 240, 160, 16]


# ---- the bf16-storage emulation of the CUDA path (oracle section "bf16-storage emulation") --------------------
@pytest.mark.parametrize("case", CASES)
def test_emulation_without_rounding_is_the_reference(golden, case):
    """With every rounding switched off, the emulated pipeline (commuted 1x1, hand-written GroupNorm and trilinear
    backward, statistics from the conv accumulators) must reproduce the real reference's golden vectors."""
    g = golden(case)
    sd = O.init_params(int(g["weight_seed"]))
    x = torch.from_numpy(g["x"])
    t = torch.from_numpy(g["target"]).float()
    for with_bce, key in ((False, "gdice::"), (True, "gboth::")):
        with O.rounding(False):
            loss, probs, logits, grads = O.train_step_bf16_emulated(sd, x, t, with_bce=with_bce)
        np.testing.assert_allclose(logits.numpy(), g["logits"], rtol=0, atol=5e-5)
        np.testing.assert_allclose(probs.numpy(), g["probs"], rtol=0, atol=1e-5)
        # Gradients: the fp32 autograd of the reference itself carries up to ~1 % error on a few tensors of this tiny
        # case (its GroupNorm backward cancels badly on 128-element groups: fp32 oracle vs float64 oracle differ by
        # 9e-3), so the tight comparison is against the oracle in float64 and the golden comparison is loose.
        sdd = {k: v.double() for k, v in sd.items()}
        _, _, g64 = O.train_step(sdd, x.double(), t.double(), with_bce=with_bce)
        for k, ref in g64.items():
            got = grads[k].double()
            assert ((got - ref).norm() / ref.norm().clamp_min(1e-30)).item() <= 2e-3, k
        for k in g.files:
            if k.startswith(key):
                ref = g[k]
                got = grads[k[len(key):]].numpy()
                assert np.linalg.norm(got - ref) <= 2e-2 * max(np.linalg.norm(ref), 1e-12), k


def test_emulation_with_rounding_stays_at_the_bf16_noise_floor(golden):
    g = golden("box16x24x32_b2")
    sd = O.init_params(int(g["weight_seed"]))
    x = torch.from_numpy(g["x"])
    t = torch.from_numpy(g["target"]).float()
    loss, probs, logits, grads = O.train_step_bf16_emulated(sd, x, t)
    ref = torch.from_numpy(g["logits"])
    err = (logits - ref).abs()
    assert 1e-3 < err.mean().item() / ref.abs().mean().item() < 2e-2        # rounded, but only at bf16 level
    assert abs(loss.item() - float(g["dice"])) < 1e-3
    rel = [np.linalg.norm(grads[k[7:]].numpy() - g[k]) / max(np.linalg.norm(g[k]), 1e-20) for k in g.files if k.startswith("gdice::")]
    assert 1e-3 < float(np.median(rel)) < 6e-2 and max(rel) < 0.2


def test_emulated_trilinear_adjoint_is_the_autograd_adjoint():
    x = torch.randn(2, 3, 4, 6, 5, dtype=torch.float64, requires_grad=True)
    y = x
    for dim in (4, 3, 2):
        y = O._up1d(y, dim)
    np.testing.assert_allclose(y.detach().numpy(), O.trilinear_x2(x.detach()).numpy(), atol=1e-12)
    gy = torch.randn_like(y)
    (gx,) = torch.autograd.grad(y, x, gy)
    with O.rounding(False):
        mine = O._up1d_adjoint(O._up1d_adjoint(O._up1d_adjoint(gy, 4), 3), 2)
    np.testing.assert_allclose(mine.numpy(), gx.numpy(), atol=1e-12)


def test_emulated_groupnorm_backward_is_the_autograd_backward():
    torch.manual_seed(0)
    c = torch.randn(2, 16, 3, 4, 5, dtype=torch.float64, requires_grad=True)
    gamma = torch.randn(16, dtype=torch.float64, requires_grad=True)
    beta = torch.randn(16, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.leaky_relu(torch.nn.functional.group_norm(c, 8, gamma, beta, O.GN_EPS), O.LRELU_SLOPE)
    gy = torch.randn_like(y)
    gc, gg, gb = torch.autograd.grad(y, (c, gamma, beta), gy)
    with O.rounding(False):
        mean, rstd = O.gn_stats_emulated(c.detach())
        dx, dg, db = O.gn_backward_emulated(c.detach(), mean.double(), rstd.double(), gamma.detach(), beta.detach(), gy, True)
    np.testing.assert_allclose(dx.numpy(), gc.numpy(), atol=2e-6)
    np.testing.assert_allclose(dg.numpy(), gg.numpy(), rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(db.numpy(), gb.numpy(), rtol=2e-6, atol=2e-6)
