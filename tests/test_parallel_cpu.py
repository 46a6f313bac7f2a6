"""CPU, world_size 2 over gloo: the host-side logic of the data-parallel path
(bucketed gradient all-reduce store, Dice-sum coupling, volume sharding)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import resunet_oracle as O


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from brats2019_b200.parallel import BucketedAllReduce, shard_volumes
        # ---- bucketed store: views of one flat buffer, reduced in >= min_bucket chunks ----
        shapes = [("a", (3, 5)), ("b", (7,)), ("c", (2, 2, 2)), ("d", (33,)), ("e", (1,))]
        flat = torch.full((sum((torch.Size(s).numel() + 3) // 4 * 4 for _, s in shapes),), float("nan"))
        flat.zero_()
        store = BucketedAllReduce(flat, dist.group.WORLD, min_bucket_elems=20)
        for i, (n, s) in enumerate(shapes):
            g = store.new(n, torch.empty(s))
            g.fill_(float((rank + 1) * (i + 1)))
            store.mark()
        store.finish()
        ok = True
        for i, (n, s) in enumerate(shapes):
            want = float(sum((r + 1) * (i + 1) for r in range(world)))
            ok &= bool((store.grads[n] == want).all()) and tuple(store.grads[n].shape) == s
        # buckets partition [0, cursor) without overlap, each (but the last) >= min_bucket
        b = store.buckets
        ok &= b[0][0] == 0 and b[-1][1] == store.cursor and all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))
        ok &= all(e - s_ >= 20 for s_, e in b[:-1]) and len(b) >= 2
        # ---- Dice coupling: all-reduced per-shard sums give the global-batch loss (loss.py:114-115) ----
        g = torch.Generator().manual_seed(0)
        p = torch.rand(4, 3, 4, 4, 4, generator=g)
        t = (torch.rand(4, 3, 4, 4, 4, generator=g) > 0.6).float()
        lo, hi = rank * 2, rank * 2 + 2
        sums = torch.zeros(8)
        sums[:3] = (p[lo:hi] * t[lo:hi]).sum(dim=(0, 2, 3, 4))
        sums[4:7] = (p[lo:hi] ** 2 + t[lo:hi]).sum(dim=(0, 2, 3, 4))
        dist.all_reduce(sums)
        loss = 1 - (2 * (sums[:3] + 1e-6) / (sums[4:7] + 2e-6)).mean()
        ok &= abs(loss.item() - O.dice_loss_joint([p], [t]).item()) < 1e-6
        # per-rank gradient of the GLOBAL loss w.r.t. local probabilities == slice of the global gradient
        gl = O.dice_loss_grad_closed_form(p, t)[lo:hi]
        I, U = sums[:3] + 1e-6, sums[4:7] + 2e-6
        mine = -(2.0 / 3) * (t[lo:hi] * U.view(1, 3, 1, 1, 1) - 2 * p[lo:hi] * I.view(1, 3, 1, 1, 1)) / U.view(1, 3, 1, 1, 1) ** 2
        ok &= torch.allclose(mine, gl, atol=1e-7)
        # ---- volume sharding ----
        shards = [list(shard_volumes(64, r, 8)) for r in range(8)]
        ok &= sorted(sum(shards, [])) == list(range(64)) and all(len(s_) == 8 for s_ in shards)
        ok &= sorted(sum([list(shard_volumes(5, r, 4)) for r in range(4)], [])) == list(range(5))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_flat_grad_layout_covers_live_parameters():
    import brats2019_b200 as B
    from brats2019_b200.parallel import flat_grad_elems
    m = B.UNet(**B.DEFAULT_CFG)
    live = m.live_parameters()
    assert len(live) == 86 and sum(p.numel() for _, p in live) == 4509939
    assert flat_grad_elems(m) >= 4509939
