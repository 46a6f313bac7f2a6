"""reference checkpoints (pickled DataParallel(model.UNet) + train.TrainingState,
train.py:320-324) load into the drop-in classes without the reference sources, and are written back in
the same layout.  The fixture was written by the real reference code: oracle/gen_ref_checkpoint.py."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR

CKPT = os.path.join(GOLDEN_DIR, "ref_checkpoint_tiny.pth")


def _load():
    assert "/root/reference" not in sys.path and "model" not in sys.modules
    import contextlib
    from brats2019_b200 import checkpoint as C
    with contextlib.redirect_stdout(sys.stderr):
        return C, C.load_reference_checkpoint(CKPT)


def test_reference_checkpoint_loads_without_reference_sources():
    C, (net, state) = _load()
    import brats2019_b200 as B
    assert isinstance(net, B.UNet) and net.depth == 2 and list(net.number_of_channels) == [16, 32]
    assert isinstance(state, C.TrainingState)
    assert state.epoch == 7 and state.global_step == 7000 and np.allclose(state.best_val, [0.9, 0.8, 0.7])
    # the weights are the pickled ones, key for key, with DataParallel's prefix stripped
    raw = torch.load(CKPT, map_location="cpu", pickle_module=C._RefPickleModule, weights_only=False)["model"]
    raw_sd = raw.state_dict()
    assert all(k.startswith("module.") for k in raw_sd)
    sd = net.state_dict()
    assert list(sd) == [k[7:] for k in raw_sd]
    for k, v in raw_sd.items():
        assert torch.equal(sd[k[7:]], v)


def test_oracle_forward_with_checkpoint_weights_matches_reference_forward():
    from oracle import resunet_oracle as O
    _, (net, _) = _load()
    ref = np.load(os.path.join(GOLDEN_DIR, "ref_checkpoint_tiny_forward.npz"))
    cfg = dict(depth=2, encoder_layers=[1, 2], decoder_layers=[1, 1], number_of_channels=[16, 32], number_of_outputs=3)
    p = O.unet_forward({k: v for k, v in net.state_dict().items()}, [torch.from_numpy(ref["x"])], cfg=cfg)[0]
    assert np.abs(p.numpy() - ref["probs"]).max() < 2e-5


def test_save_reference_layout_roundtrip(tmp_path):
    C, (net, state) = _load()
    out = str(tmp_path / "resaved.pth")
    C.save_reference_layout(out, net, state)
    s = torch.load(out, map_location="cpu", weights_only=False)
    assert set(s) == {"state", "model"}
    keys = list(s["model"].state_dict())
    assert keys and all(k.startswith("module.") for k in keys)          # export_onnx_group_norm.py:28-31 strips this
    net2, state2 = C.load_reference_checkpoint(out)
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k
    assert state2.global_step == 7000
    # Trainer._load with an existing model (train.py:331-333)
    import brats2019_b200 as B
    fresh = B.UNet(depth=2, encoder_layers=[1, 2], decoder_layers=[1, 1], number_of_channels=[16, 32], number_of_outputs=3)
    C.load_state_dict_into(C.ModulePrefix(fresh), out)
    assert torch.equal(fresh.conv_output.weight, net.conv_output.weight)


@pytest.mark.gpu
def test_checkpoint_forward_on_gpu_matches_reference_forward():
    _, (net, _) = _load()
    ref = np.load(os.path.join(GOLDEN_DIR, "ref_checkpoint_tiny_forward.npz"))
    net = net.cuda().eval()
    with torch.no_grad():
        p = net([torch.from_numpy(ref["x"]).cuda()])[0].cpu().numpy()
    # bf16 storage vs the reference's fp32 forward: the tolerance of tests/test_model_gpu.py (probs <= 0.08)
    assert np.abs(p - ref["probs"]).max() < 0.08
    assert ((p > 0.5) == (ref["probs"] > 0.5)).mean() > 0.98
