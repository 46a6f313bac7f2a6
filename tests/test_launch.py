"""the launcher that replaces main.py (one process per GPU, the reference's recipe)."""
import math
import os

import pytest
import torch


def test_weight_init_matches_reference_distributions():
    """weight_init.py:22-27: Conv3d weights kaiming-normal(a=1e-2, leaky_relu), Conv3d bias N(0,1)."""
    from brats2019_b200.launch import weight_init
    m = torch.nn.Conv3d(32, 64, 3, bias=True)
    weight_init(m, torch.Generator().manual_seed(1))
    want = torch.nn.init.calculate_gain("leaky_relu", 1e-2) / math.sqrt(32 * 27)
    assert abs(m.weight.std().item() / want - 1) < 0.02
    assert abs(m.bias.std().item() - 1) < 0.3
    gn = torch.nn.GroupNorm(8, 16)
    weight_init(gn)
    assert torch.equal(gn.weight, torch.ones(16)) and torch.equal(gn.bias, torch.zeros(16))


def test_cli_refuses_cpu():
    from brats2019_b200 import launch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU path"):
        launch.main(["--steps", "1"])


@pytest.mark.gpu
def test_launcher_trains_and_writes_reference_checkpoint(tmp_path):
    from brats2019_b200 import checkpoint as C
    from brats2019_b200 import launch
    from oracle import resunet_oracle as O
    dev = torch.device("cuda", 0)
    launch.seed_everything(1337)
    net, criteria, group = launch.build(dev, distributed=False)
    assert group is None
    sd0 = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    batches = list(launch.synthetic_batches(1, (32, 32, 32), 4, 100, dev))
    lines = []
    hist, state = launch.train(net, criteria, batches, log_every=1, log=lines.append)
    torch.cuda.synchronize()
    assert state.global_step == 4 and len(hist) == 4 and all(math.isfinite(v) for h in hist for v in h)
    # first step against the CPU oracle on the same weights and batch: [Dice, BCE(bg 1e-2)] (main.py:126-128)
    x, t = batches[0][0][0].cpu(), batches[0][1][0].cpu()
    probs = O.unet_forward(sd0, [x])
    dice = O.dice_loss_joint(probs, [t]).item()
    bce = O.bce_loss(probs, [t], bg_weight=1e-2).item()
    assert abs(hist[0][0] - dice) < 2e-3 and abs(hist[0][1] - bce) < 0.02 * abs(bce) + 2e-3
    # Adam moved the weights, and the optimizer state is recorded like train.py:315
    assert not torch.equal(net.state_dict()["conv_output.weight"].cpu(), sd0["conv_output.weight"])
    assert state.optimizer_state is not None
    out = str(tmp_path / "last_model.pth")
    C.save_reference_layout(out, net, state)
    net2, state2 = C.load_reference_checkpoint(out)
    assert state2.global_step == 4
    assert torch.equal(net2.state_dict()["conv_input.weight"], net.state_dict()["conv_input.weight"].cpu())
