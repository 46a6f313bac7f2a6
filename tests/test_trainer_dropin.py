"""Driver parity: the reference's own `train.Trainer` - `_train_one_epoch` (train.py:178-241),
`predict` (:129-143), `predict_tiled` (:145-176, one 192^3 tile per 48^3 centre) and `_save` (:320-324) - drives the
drop-in `brats2019_b200.UNet` unchanged.  Where the reference checkout exists the unmodified class is imported
(tests/trainer_harness.py, three stub modules for its absent third-party imports); on the GPU box, where it does
not, the same call sequences restated line by line are used.  Which one ran is printed."""
import os

import pytest
import torch

from oracle import resunet_oracle as O
import trainer_harness as TH


def _net():
    import brats2019_b200 as B
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(O.init_params(1337))
    return m


def test_reference_trainer_accepts_the_dropin_on_cpu(tmp_path):
    """CPU pre-flight with the REAL reference Trainer: construction (train.py:27-52 asserts an nn.Module), `_save`
    (pickles the whole module object, train.py:320-324) and `predict` reaching our forward with the list-of-tensors
    call shape - where it must refuse to run on the CPU instead of falling back."""
    if TH.reference_dir() is None:
        pytest.skip("reference checkout not present")
    Trainer, _, kind = TH.load()
    assert kind == "reference"
    import brats2019_b200 as B
    from brats2019_b200 import checkpoint as C
    m = _net()
    tr = Trainer(name="run0", models_root=str(tmp_path), model=m, connect_tb=True)
    tr._save("last_model")
    path = os.path.join(str(tmp_path), "run0", "run0last_model.pth")
    assert os.path.exists(path)
    m2, state = C.load_reference_checkpoint(path)
    assert isinstance(m2, B.UNet) and state.global_step == 0
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    tr.state.cuda = False
    with pytest.raises(RuntimeError, match="no CPU"):
        tr.predict([[torch.zeros(1, 4, 16, 16, 16)]])


@pytest.mark.gpu
def test_trainer_drives_the_dropin(tmp_path):
    import brats2019_b200 as B
    from brats2019_b200 import checkpoint as C
    from brats2019_b200.optim import FusedAdam
    Trainer, Dice, kind = TH.load()
    print("driver:", kind, Trainer.__module__)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    m = _net()
    tr = Trainer(name="run0", models_root=str(tmp_path), model=m, connect_tb=True)
    tr.cuda()
    # ---- _train_one_epoch: criterion list, optimizer and scheduler as main.py:126-142 builds them ----
    g = torch.Generator().manual_seed(5)
    loader = [([torch.randn(2, 4, 32, 32, 32, generator=g)], [(torch.rand(2, 3, 32, 32, 32, generator=g) > 0.7).float()])
              for _ in range(3)]
    criterion = [B.Dice_loss_joint(index=0, priority=1), B.BCE_Loss(index=0, bg_weight=1e-2)]
    opt = FusedAdam(m.parameters(), lr=2e-5, weight_decay=1e-6, amsgrad=True, model=m)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=16000, gamma=0.5)
    metric = Dice(name="Dice", input_index=0, target_index=0, classes=4)
    results = {"Dice": []}
    w0 = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
    step = tr._train_one_epoch(criterion=criterion, optimizer=opt, training_data_loader=loader, train_metrics=[metric],
                               train_metrics_results=results, epoch=0, global_step=0, scheduler=sched)
    assert step == 3 and len(results["Dice"]) == 1 and tr.state.optimizer_state is not None
    # first step against the oracle on the initial weights: [Dice, BCE] values the trainer logged
    x0, t0 = loader[0][0][0], loader[0][1][0]
    probs = O.unet_forward({k: v.cuda() for k, v in w0.items()}, [x0.cuda()])
    dice, bce = O.dice_loss_joint(probs, [t0.cuda()]).item(), O.bce_loss(probs, [t0.cuda()], bg_weight=1e-2).item()
    logged = [v for tag, v, s in tr.tb_writer.scalars if tag.startswith("loss/") and s == 0] if kind == "reference" else tr.logged[0]
    assert abs(logged[0] - dice) < 1e-3 and abs(logged[1] - bce) < 0.02 * abs(bce) + 2e-3, (logged, dice, bce)
    assert not torch.equal(m.state_dict()["conv_output.weight"].cpu(), w0["conv_output.weight"])      # Adam stepped
    # ---- predict: eval / no_grad forward of the current weights, list in, list out ----
    sd_now = {k: v.detach() for k, v in m.state_dict().items()}
    xp = torch.randn(1, 4, 48, 64, 32, generator=g)
    out = tr.predict([[xp]])
    assert isinstance(out, list) and out[0].shape == (1, 3, 48, 64, 32) and out[0].is_cuda and not m.training
    ref = O.unet_forward(sd_now, [xp.cuda()])[0]
    assert (out[0] - ref).abs().max().item() < 0.08
    # ---- predict_tiled: 192^3 tiles with a 48^3 centre (train.py:154-156), two tiles for a 96 x 48 x 48 volume ----
    xv = torch.randn(1, 4, 96, 48, 48, generator=g)
    tiled = tr.predict_tiled([[xv]], (1, 3, 96, 48, 48))[0]
    with torch.no_grad():
        ref_t = O.predict_tiled(lambda t: O.unet_forward(sd_now, [t.cuda()])[0].cpu(), xv, (1, 3, 96, 48, 48))
    # a 192^3 tile around a 96 x 48 x 48 volume is 98 % zero padding: GroupNorm normalises almost-constant tensors and single
    # voxels can swing (max |dp| ~0.2); the bulk statistics are the meaningful ones
    err = (tiled - ref_t).abs().flatten()
    assert tiled.shape == ref_t.shape and err.mean().item() < 5e-3 and err.kthvalue(int(0.999 * err.numel())).values.item() < 0.15
    a, b = tiled > 0.5, ref_t > 0.5
    assert 2.0 * (a & b).sum().item() / max(1, a.sum().item() + b.sum().item()) > 0.98
    # ---- _save: the whole module is pickled; our reader restores class and weights ----
    tr._save("last_model")
    m2, state = C.load_reference_checkpoint(os.path.join(str(tmp_path), "run0", "run0last_model.pth"))
    assert isinstance(m2, B.UNet)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a.cpu(), b.cpu()), k
    del m, m2
    torch.cuda.empty_cache()
