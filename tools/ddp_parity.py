"""Multi-GPU parity (run under torchrun, one rank per GPU):

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_parity.py

Each rank takes its shard of a global batch; the Dice sums and the gradients are all-reduced
(brats2019_b200.parallel).  Rank 0 then runs the SAME global batch alone and compares loss and
every gradient: SUM-reduced data-parallel gradients must equal the single-process gradients on the
concatenated batch (the reference's DataParallel semantics, main.py:61)."""
import contextlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import brats2019_b200 as B  # noqa: E402
from brats2019_b200.parallel import DistributedUNet  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    per = 2
    S = (32, 32, 48)
    g = torch.Generator().manual_seed(5)
    X = torch.randn(world * per, 4, *S, generator=g)
    T = (torch.rand(world * per, 3, *S, generator=g) > 0.7).float()
    torch.manual_seed(100 + rank)          # different init per rank: the wrapper must broadcast rank 0's weights
    with contextlib.redirect_stdout(sys.stderr):
        m = B.UNet(**B.DEFAULT_CFG)
    for p in m.parameters():
        if p.dim() > 1:
            torch.nn.init.kaiming_normal_(p, a=1e-2)
    m = m.to(dev).train()
    net = DistributedUNet(m, min_bucket_elems=1 << 18)
    crit = B.Dice_loss_joint()
    crit.process_group = net.process_group
    x, t = X[rank * per:(rank + 1) * per].to(dev), T[rank * per:(rank + 1) * per].to(dev)
    loss = crit(net([x]), [t])
    loss.backward()
    torch.cuda.synchronize()
    grads = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
    buckets = list(net.last_buckets)
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, loss.detach().reshape(1))
    ok = True
    if rank == 0:
        same_loss = all(abs(l.item() - losses[0].item()) < 1e-7 for l in losses)
        print("buckets:", len(buckets), [e - s for s, e in buckets], "loss per rank:", [round(l.item(), 6) for l in losses])
        ok &= same_loss and len(buckets) >= 2
        # single-process run on the whole batch with the same (rank-0) weights
        crit1 = B.Dice_loss_joint()
        m.zero_grad(set_to_none=True)
        m._grad_store_factory = None
        loss1 = crit1(m([X.to(dev)]), [T.to(dev)])
        loss1.backward()
        torch.cuda.synchronize()
        worst = 0.0
        for n, p in m.named_parameters():
            if p.grad is None:
                assert n not in grads
                continue
            rel = ((grads[n] - p.grad).norm() / p.grad.norm().clamp_min(1e-20)).item()
            worst = max(worst, rel)
        print("loss ddp %.7f single %.7f | worst gradient rel-L2 difference %.3e" % (losses[0].item(), loss1.item(), worst))
        # the two runs differ only by fp32 summation order (wgrad K-splits, GN partials) and bf16
        # re-rounding of identical values: agreement must be far tighter than the bf16 tolerance
        ok &= abs(losses[0].item() - loss1.item()) < 1e-5 and worst < 2e-2
        print("DDP PARITY", "OK" if ok else "FAIL")
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
