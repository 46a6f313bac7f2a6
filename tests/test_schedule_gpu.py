"""The launch STRUCTURE of a training step must not change a single bit of its result.

Round 2b reordered the backward pass (weight gradients of the level boundaries on the side stream, deferred launches,
`B200_BWD_SCHED`), made the conv kernels programmatic dependents of their stream predecessors (`B200_PDL`: their
prologues and weight copies run before `griddepcontrol.wait`) and overlapped the weight re-pack with the input
conversion.  All of that only moves kernels in time: the loss, the probabilities and all 86 gradients of a step must be
bit-identical with every switch setting - a race (a gradient read before its producer finished, a buffer overwritten
under a side-stream reader, weights copied before they were packed) shows up here as a difference.  Each setting runs the
step three times, eagerly and as a CUDA-graph replay, after a weight update (so the re-pack path is on the clock)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

SETTINGS = [
    {},                                                     # defaults: new order, PDL level 1
    {"B200_BWD_SCHED": "0"},                                # round-2a launch order
    {"B200_PDL": "0"},                                      # ordinary launches
    {"B200_PDL": "2"},                                      # every kernel a programmatic dependent
    {"B200_NO_WGRAD_OVERLAP": "1"},                         # one stream
    {"B200_PDL_WGRAD": "1"},
    {"B200_PDL_EW_MAX": "600"},                             # memory-bound kernels of the deep levels (few CTAs) as dependents
]


def _run(env, shape, graphed):
    import brats2019_b200 as B
    from brats2019_b200.graphs import GraphedTrainStep
    from brats2019_b200.optim import FusedAdam
    saved = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        torch.manual_seed(3)
        m = B.UNet(**B.DEFAULT_CFG).cuda().train()
        crit = B.Dice_loss_joint()
        g = torch.Generator(device="cuda").manual_seed(11)
        x = torch.randn(*shape, device="cuda", generator=g)
        t = (torch.rand(shape[0], 3, *shape[2:], device="cuda", generator=g) > 0.7).float()
        out = []
        if graphed:
            opt = FusedAdam(m.parameters(), lr=1e-3, weight_decay=1e-6, amsgrad=True, model=m)
            step = GraphedTrainStep(m, crit, opt, x, t, warmup=2)
            for _ in range(3):
                loss = step()
            torch.cuda.synchronize()
            out.append(loss.detach().clone())
            out += [p.detach().clone() for p in m.parameters()]          # three optimizer steps deep
        else:
            for it in range(3):
                for p in m.parameters():
                    p.grad = None
                probs = m([x])
                loss = crit(probs, [t])
                loss.backward()
                with torch.no_grad():                                     # touch the weights: the next forward re-packs
                    for p in m.parameters():
                        if p.grad is not None:
                            p.add_(p.grad, alpha=-1e-2)
            torch.cuda.synchronize()
            out.append(loss.detach().clone())
            out.append(probs[0].detach().clone())
            out += [p.grad.detach().clone() for p in m.parameters() if p.grad is not None]
        return out
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


@pytest.mark.parametrize("graphed", [False, True])
def test_every_launch_structure_gives_the_same_bits(graphed):
    shape = (2, 4, 32, 48, 64)
    ref = _run(SETTINGS[0], shape, graphed)
    again = _run(SETTINGS[0], shape, graphed)
    assert all(torch.equal(a, b) for a, b in zip(ref, again)), "the default path is not repeatable"
    for env in SETTINGS[1:]:
        got = _run(env, shape, graphed)
        assert len(got) == len(ref)
        bad = [i for i, (a, b) in enumerate(zip(ref, got)) if not torch.equal(a, b)]
        assert not bad, "%s changes %d of %d tensors (first: #%d, max diff %.3g)" % (
            env, len(bad), len(ref), bad[0], (ref[bad[0]].float() - got[bad[0]].float()).abs().max().item())


def test_config3_shape_overlapped_step_equals_the_fully_serial_one():
    """BASELINE config 3 (batch 2 x 4x128^3), CUDA-graph replays: the default step (two streams, deferred launches,
    programmatic dependents) against the same step with one stream, the old order and ordinary launches."""
    shape = (2, 4, 128, 128, 128)
    ref = _run({}, shape, True)
    got = _run({"B200_PDL": "0", "B200_BWD_SCHED": "0", "B200_NO_WGRAD_OVERLAP": "1"}, shape, True)
    bad = [i for i, (a, b) in enumerate(zip(ref, got)) if not torch.equal(a, b)]
    assert not bad, "%d of %d tensors differ" % (len(bad), len(ref))
