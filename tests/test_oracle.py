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
