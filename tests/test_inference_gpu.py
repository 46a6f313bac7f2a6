"""GPU: the inference procedures around the model (pad, 4-flip TTA, sliding window,
label painting) against the CPU oracle's restatement of test.py:92-159 / train.py:145-176, plus
size-independent properties at the BraTS inference shape."""
import contextlib
import sys

import numpy as np
import pytest
import torch

from oracle import resunet_oracle as O

pytestmark = pytest.mark.gpu


def _model():
    import brats2019_b200 as B
    with contextlib.redirect_stdout(sys.stderr):
        m = B.UNet(**B.DEFAULT_CFG)
    sd = O.init_params(1337)
    m.load_state_dict(sd)
    return m.cuda().eval(), sd


def _mask_agreement(p, ref):
    return ((p > 0.5) == (ref > 0.5)).float().mean().item()


def test_tta_matches_oracle():
    from brats2019_b200 import inference as I
    m, sd = _model()
    g = torch.Generator().manual_seed(21)
    raw = torch.randn(1, 4, 13, 24, 30, generator=g)                 # 13, 30 are not multiples of 16 (test.py:93)
    xp, left, right = I.pad_to_multiple(raw, 16)
    assert xp.shape[2:] == (16, 32, 32)
    ref = O.tta_predict(sd, xp)
    with torch.no_grad():
        got = I.predict_tta(m, xp.cuda()).cpu()
    assert got.shape == ref.shape
    assert (got - ref).abs().max().item() < 0.08                      # tolerance of tests/test_model_gpu.py (bf16 vs fp32)
    assert _mask_agreement(got, ref) > 0.98
    # un-padding returns the original extent
    assert I.unpad(got, left, right).shape[2:] == raw.shape[2:]


def test_sliding_window_matches_oracle_procedure():
    """Same tile geometry rules as Trainer.predict_tiled, small numbers: tile 32, centre 16, border 8."""
    from brats2019_b200 import inference as I
    m, sd = _model()
    g = torch.Generator().manual_seed(22)
    x = torch.randn(1, 4, 40, 24, 32, generator=g)                    # 40/16 -> 3 tiles with a ragged last centre
    tile, center, border = (32, 32, 32), (16, 16, 16), (8, 8, 8)
    with torch.no_grad():
        ref = O.predict_tiled(lambda t: O.unet_forward(sd, [t])[0], x, (1, 3, 40, 24, 32), tile, center, border)
        for bt in (1, 4):
            got = I.predict_tiled(m, x.cuda(), 3, tile, center, border, batch_tiles=bt).cpu()
            assert (got - ref).abs().max().item() < 0.08, bt
            assert _mask_agreement(got, ref) > 0.98, bt


def test_label_painting_matches_reference_rule():
    """test.py:144-159: WT -> 2, TC -> 1, ET -> 4 only when more than 32 ET voxels."""
    from brats2019_b200 import inference as I
    g = torch.Generator().manual_seed(23)
    probs = torch.rand(2, 3, 8, 12, 10, generator=g)
    probs[1, 2] = 0.0
    probs[1, 2, 0, 0, :5] = 0.9                                       # 5 ET voxels: below the 32-voxel rule
    got = I.paint_labels(probs.cuda()).cpu().numpy()
    want = np.zeros((2, 8, 12, 10), np.uint8)
    pn = probs.numpy()
    for b in range(2):
        want[b][pn[b, 0] > 0.5] = 2
        want[b][pn[b, 1] > 0.5] = 1
        if (pn[b, 2] > 0.5).sum() > 32:
            want[b][pn[b, 2] > 0.5] = 4
    np.testing.assert_array_equal(got, want)
    assert 4 in got[0] and 4 not in got[1]


def test_batch_independence_and_determinism_at_brats_shape():
    """GroupNorm statistics are per sample (model.py:95): a volume's result must not depend on what else is in
    the batch, and the same input must give bit-identical output run to run.  4x240x240x160 (config 2)."""
    m, _ = _model()
    g = torch.Generator().manual_seed(24)
    a = torch.randn(1, 4, 240, 240, 160, generator=g).cuda()
    b = torch.randn(1, 4, 240, 240, 160, generator=g).cuda()
    with torch.no_grad():
        pa1 = m([a])[0].clone()
        pa2 = m([a])[0].clone()
        pab = m([torch.cat([a, b])])[0]
    assert torch.equal(pa1, pa2)
    assert torch.isfinite(pab).all()
    # Batch 1 and batch 2 use different tile plans, so the fp32 GroupNorm partial sums are added in a different
    # order; a last-bit change of a mean flips bf16 roundings downstream.  The result must agree to the bf16
    # tolerance of tests/test_model_gpu.py (probs <= 0.08 max), and far tighter on average.
    diff = (pab[:1] - pa1).abs()
    assert diff.max().item() < 0.08
    assert diff.mean().item() < 2e-3
    assert _mask_agreement(pab[:1], pa1) > 0.995
    # zero padding stays inert: the padded slab of config 2 (155 -> 160) does not produce NaN/Inf
    a[..., 155:] = 0
    with torch.no_grad():
        assert torch.isfinite(m([a])[0]).all()
