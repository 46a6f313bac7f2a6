"""the fused optimizer step (b200_adam_step) against torch.optim.Adam - the optimizer the
reference builds at main.py:133-138 and steps at train.py:220 - and the CUDA-graph-replayed launcher loop."""
import copy
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _clone_params(shapes, seed):
    g = torch.Generator().manual_seed(seed)
    return [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]


@pytest.mark.parametrize("amsgrad,wd", [(True, 1e-6), (True, 0.1), (False, 0.0)])
def test_trajectory_equals_torch_adam(amsgrad, wd):
    """Ten steps on random gradients: parameters and both moment buffers equal torch's to 1e-6 (relative to the
    tensor's scale), with a StepLR(3, 0.5) schedule stepped once per batch and one parameter that never gets a
    gradient (torch skips it entirely: no weight decay, no state)."""
    from brats2019_b200.optim import FusedAdam
    shapes = [(16, 4, 3, 3, 3), (16,), (33, 7), (128, 128, 3, 3, 3), (5,)]
    pa, pb = _clone_params(shapes, 1), _clone_params(shapes, 1)
    lr = 1e-2
    ref = torch.optim.Adam(pa, lr=lr, weight_decay=wd, amsgrad=amsgrad)
    sched = torch.optim.lr_scheduler.StepLR(ref, step_size=3, gamma=0.5)
    mine = FusedAdam(pb, lr=lr, weight_decay=wd, amsgrad=amsgrad, lr_step_size=3, lr_gamma=0.5)
    g = torch.Generator().manual_seed(2)
    for it in range(10):
        for a, b in zip(pa[:-1], pb[:-1]):                   # the last parameter is "dead"
            gr = torch.randn(a.shape, generator=g).cuda() * (1.0 + it)
            a.grad, b.grad = gr.clone(), gr.clone()
        ref.step(); sched.step(); mine.step()
        ref.zero_grad(); mine.zero_grad()
    torch.cuda.synchronize()
    for i, (a, b) in enumerate(zip(pa, pb)):
        scale = a.detach().abs().max().item()
        assert (a.detach() - b.detach()).abs().max().item() <= 1e-6 * scale, (i, (a - b).abs().max().item(), scale)
    assert torch.equal(pa[-1].detach(), pb[-1].detach()) and pb[-1] not in mine.state    # untouched, no state
    for a, b in zip(pa[:-1], pb[:-1]):
        sa, sb = ref.state[a], mine.state[b]
        for k in ("exp_avg", "exp_avg_sq") + (("max_exp_avg_sq",) if amsgrad else ()):
            assert (sa[k] - sb[k]).abs().max().item() <= 2e-6 * sa[k].abs().max().item(), k
        assert float(sb["step"]) == 10.0
    # state dict in torch's layout: loads into torch.optim.Adam and continues identically
    ref2 = torch.optim.Adam(pb, lr=lr, weight_decay=wd, amsgrad=amsgrad)
    ref2.load_state_dict(mine.state_dict())
    assert float(ref2.state[pb[0]]["step"]) == 10.0
    # ... and torch's state loads into a fresh FusedAdam (resume from a reference checkpoint, train.py:83-84)
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    mine2 = FusedAdam(pc, lr=lr, weight_decay=wd, amsgrad=amsgrad)      # torch's param_groups carry the decayed lr
    mine2.load_state_dict(ref.state_dict())
    gr = [torch.randn(a.shape, generator=g).cuda() for a in pa[:-1]]
    for a, c, q in zip(pa[:-1], pc[:-1], gr):
        a.grad, c.grad = q.clone(), q.clone()
    ref.step(); mine2.step()
    for a, c in zip(pa, pc):
        assert (a.detach() - c.detach()).abs().max().item() <= 1e-6 * a.detach().abs().max().item()


def _net(seed=1337):
    import brats2019_b200 as B
    from oracle import resunet_oracle as O
    m = B.UNet(**B.DEFAULT_CFG)
    m.load_state_dict(O.init_params(seed))
    return m.cuda().train()


def test_bound_to_the_model_equals_torch_adam_and_copies_nothing():
    """The recipe of main.py:133-138 on the real model: after the first step the engine writes gradients straight
    into the optimizer's flat buffer (p.grad aliases it) and ten steps give the weights torch.optim.Adam gives."""
    import brats2019_b200 as B
    from brats2019_b200.optim import FusedAdam
    ma, mb = _net(), _net()
    ref = torch.optim.Adam(ma.parameters(), lr=1e-3, weight_decay=1e-6, amsgrad=True)
    mine = FusedAdam(mb.parameters(), lr=1e-3, weight_decay=1e-6, amsgrad=True, model=mb)
    crit = B.Dice_loss_joint()
    g = torch.Generator().manual_seed(3)
    la, lb = [], []
    w0 = {n: p.detach().clone() for n, p in ma.named_parameters()}
    for it in range(10):
        x = torch.randn(2, 4, 16, 16, 16, generator=g).cuda()
        t = (torch.rand(2, 3, 16, 16, 16, generator=g) > 0.7).float().cuda()
        for m, o, ls in ((ma, ref, la), (mb, mine, lb)):
            o.zero_grad(set_to_none=True)
            loss = crit(m([x]), [t])
            loss.backward()
            o.step()
            ls.append(loss.item())
        if it == 0:      # identical weights and deterministic kernels give identical gradients: the first update is
            for (n, a), (_, b) in zip(ma.named_parameters(), mb.named_parameters()):      # torch's to float rounding
                assert (a.detach() - b.detach()).abs().max().item() <= 1e-6 * max(a.detach().abs().max().item(), 1e-3), n
        if it >= 1:
            L = mine._layout
            lo, hi = L["G"].data_ptr(), L["G"].data_ptr() + 4 * L["total"]
            assert all(lo <= p.grad.data_ptr() < hi for p in L["params"]), "gradients were copied, not written in place"
    dead = set(mb.dead_parameter_names())
    assert len(mine._layout["params"]) == 86 and all(mb.get_parameter(n).grad is None for n in dead)
    # From the second step on a 1e-7 weight difference flips bf16 roundings in the forward pass, the gradients of
    # the two runs differ at the bf16 noise level and Adam's normalised update (+-lr per element wherever the
    # gradient is noise) amplifies that: the trajectories stay close relative to the distance travelled, not bitwise.
    num = sum((a.detach() - b.detach()).pow(2).sum().item() for a, b in zip(ma.parameters(), mb.parameters()))
    den = sum((a.detach() - w0[n]).pow(2).sum().item() for n, a in ma.named_parameters())
    print("10 steps: |w_fused - w_torch| / |w_torch - w_0| = %.4f; losses" % math.sqrt(num / den), la, lb)
    assert math.sqrt(num / den) < 0.15
    assert max(abs(a - b) for a, b in zip(la, lb)) < 5e-3, (la, lb)
    assert lb[-1] < lb[0]
    # eval forward after the last step sees the updated weights (packed bf16 images refreshed)
    from oracle import resunet_oracle as O
    mb.eval()
    with torch.no_grad():
        x = torch.randn(1, 4, 16, 16, 16, generator=g).cuda()
        ref_probs = O.unet_forward({k: v.detach() for k, v in mb.state_dict().items()}, [x])[0]
        assert (mb([x])[0] - ref_probs).abs().max().item() < 0.08


def test_graph_replayed_launcher_equals_eager_launcher():
    """launch.train with one CUDA-graph replay per batch vs the same loop launched kernel by kernel."""
    from brats2019_b200 import launch
    dev = torch.device("cuda", 0)
    hists, finals = [], []
    for graph in (False, True):
        launch.seed_everything(1337)
        net, criteria, _ = launch.build(dev, distributed=False)
        batches = list(launch.synthetic_batches(1, (16, 16, 16), 8, 100, dev))
        hist, state = launch.train(net, criteria, batches, lr=1e-3, step_size=3, gamma=0.5, log_every=1,
                                   log=lambda *a: None, graph=graph)
        torch.cuda.synchronize()
        assert state.global_step == 8
        hists.append(hist)
        finals.append({k: v.detach().clone() for k, v in net.state_dict().items()})
    for a, b in zip(*hists):
        assert all(abs(u - v) < 5e-4 for u, v in zip(a, b)), (hists)
    for k in finals[0]:
        a, b = finals[0][k], finals[1][k]
        assert (a - b).abs().max().item() <= 5e-5 * max(a.abs().max().item(), 1e-3) + 5e-6, k


def test_two_forwards_then_backward_is_an_error_not_a_wrong_gradient():
    """the autograd node owns (plan, generation); a second training forward of the same shape overwrites
    the activations, and the first node's backward must raise instead of using them."""
    import brats2019_b200 as B
    m = _net()
    crit = B.Dice_loss_joint()
    x1 = torch.randn(1, 4, 16, 16, 16, device="cuda")
    x2 = torch.randn(1, 4, 16, 16, 16, device="cuda")
    t = (torch.rand(1, 3, 16, 16, 16, device="cuda") > 0.7).float()
    l1 = crit(m([x1]), [t])
    l2 = crit(m([x2]), [t])
    with pytest.raises(RuntimeError, match="overwritten by a later forward"):
        (l1 + l2).backward()
    # different shapes do not collide: each node finds its own plan
    m.zero_grad()
    xa = torch.randn(1, 4, 16, 16, 16, device="cuda")
    xb = torch.randn(1, 4, 16, 16, 32, device="cuda")
    tb = (torch.rand(1, 3, 16, 16, 32, device="cuda") > 0.7).float()
    la, lb = crit(m([xa]), [t]), crit(m([xb]), [tb])
    (la + lb).backward()
    ga = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    m.zero_grad()
    crit(m([xa]), [t]).backward()
    crit(m([xb]), [tb]).backward()
    for n, p in m.named_parameters():
        if p.grad is not None:
            assert torch.allclose(p.grad, ga[n], rtol=1e-4, atol=1e-7), n


def test_data_mutation_paths_refresh_the_packed_weights():
    """`.data` edits do not bump version counters; the paths the reference uses (apply(weight_init),
    load_state_dict, .cuda()) and an explicit invalidate must all be seen by the next forward."""
    from oracle import resunet_oracle as O
    m = _net(1).eval()
    x = torch.randn(1, 4, 16, 16, 16, device="cuda")
    with torch.no_grad():
        p0 = m([x])[0].clone()

        def reinit(mod):                                   # what weight_init.py:22-27 does
            if isinstance(mod, torch.nn.Conv3d):
                torch.nn.init.kaiming_normal_(mod.weight.data, a=1e-2)
        m.apply(reinit)
        p1 = m([x])[0].clone()
        assert (p1 - p0).abs().max().item() > 1e-3
        m.load_state_dict(O.init_params(1))
        assert torch.equal(m([x])[0], p0)
        for p in m.parameters():
            p.data.mul_(1.0)                                # no-op edit through .data ...
        m.conv_output.bias.data.add_(5.0)                   # ... and a real one
        m.invalidate_packed_weights()
        assert (m([x])[0] - p0).abs().max().item() > 1e-3
