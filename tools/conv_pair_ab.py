"""conv3 128 -> 128 at the deep level: CTA-pair kernel (conv_pair.cuh) vs the one-CTA generic kernel, 20 launches per graph.
    python tools/conv_pair_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

dev = "cuda"


def timed(fn, reps=20):
    gr = torch.cuda.CUDAGraph()
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


with torch.cuda.stream(torch.cuda.Stream()):
    for (B, s, Cc) in ((2, 16, 128), (8, 16, 128), (2, 24, 128)):
        res = {}
        outs = {}
        for pair in ("0", "1"):
            os.environ["B200_CONV_PAIR"] = pair
            torch.manual_seed(0)
            x = ops.act_zeros(B, s, s, s, Cc, dev)
            x.interior().copy_(torch.randn(Cc // 8, B, s, s, s, 8, device=dev).to(torch.bfloat16))
            y = ops.act_zeros(B, s, s, s, Cc, dev)
            w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
            desc = ops.conv_desc(ops.MODE_K3, B, s, s, s, Cc, Cc)
            pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
            st = torch.empty(ops.conv_ctas(desc) * B * 16, device=dev)
            res[pair] = (timed(lambda: ops.conv_run(desc, x, pk, y, stats=st)), ops.conv_ctas(desc))
            outs[pair] = (y.t.clone(), st.view(-1, B, 16).sum(0).clone())
        os.environ.pop("B200_CONV_PAIR")
        same = torch.equal(outs["0"][0], outs["1"][0])
        dst = (outs["0"][1] - outs["1"][1]).abs().max().item() / outs["0"][1].abs().max().item()
        print("conv3 %d->%d @ %dx%d^3: one CTA %.1f us (ctas %d) | pair %.1f us (ctas %d) | outputs identical: %s, stats rel diff %.2g"
              % (Cc, Cc, B, s, res["0"][0], res["0"][1], res["1"][0], res["1"][1], same, dst), flush=True)
