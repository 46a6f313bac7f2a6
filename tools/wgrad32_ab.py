"""wgrad 3x3x3 32 <-> 32 at 2 x 64^3: line-marching kernel (wgrad_line_kernel<4>) vs the linear-row kernel, 20 launches
per CUDA graph.   python tools/wgrad32_ab.py"""
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
    for (B, s, Cc) in ((2, 64, 32), (2, 128, 16), (8, 64, 32)):
        x = ops.act_zeros(B, s, s, s, Cc, dev); dy = ops.act_zeros(B, s, s, s, Cc, dev)
        for a in (x, dy):
            a.interior().copy_(torch.randn(Cc // 8, B, s, s, s, 8, device=dev).to(torch.bfloat16))
        wd = ops.wgrad_desc(0, B, s, s, s, Cc, Cc)
        ws = ops.wgrad_workspace(wd, dev)
        fl = 2.0 * B * s ** 3 * Cc * Cc * 27
        res = {}
        for form in ("line", "linear"):
            os.environ["B200_NO_WGRAD_LINE"] = "0" if form == "line" else "1"
            os.environ["B200_WGRAD_LINE32"] = "1" if form == "line" else "0"
            g = torch.empty(Cc, Cc, 3, 3, 3, device=dev)
            t = timed(lambda: ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws))
            res[form] = (t, g.clone())
        os.environ.pop("B200_NO_WGRAD_LINE")
        rel = ((res["line"][1] - res["linear"][1]).norm() / res["linear"][1].norm()).item()
        print("wgrad3 %dx%d @ %dx%d^3: line %.1f us (%.0f TFLOP/s) | linear %.1f us (%.0f TFLOP/s) | rel diff %.2g"
              % (Cc, Cc, B, s, res["line"][0], fl / res["line"][0] / 1e6, res["linear"][0], fl / res["linear"][0] / 1e6, rel), flush=True)
