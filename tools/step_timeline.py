"""Timeline of one steady-state training step (CUDA-graph replay): every kernel's start, duration and stream from
CUPTI (torch.profiler), written as CSV for offline analysis of the main-stream / side-stream overlap.

    python tools/step_timeline.py [B] [S] [out.csv] [--all]
--all keeps both profiled replays and the memset / memcpy nodes too (what sits between two replays, what precedes the
first kernel of a step).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import brats2019_b200 as B  # noqa: E402
from brats2019_b200.graphs import GraphedTrainStep  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/step_timeline.csv"
torch.manual_seed(0)
dev = torch.device("cuda")
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    m = B.UNet(**B.DEFAULT_CFG).cuda().train()
    crit = B.Dice_loss_joint()
    from brats2019_b200.optim import FusedAdam  # noqa: E402
    opt = FusedAdam(m.parameters(), lr=2e-5, weight_decay=1e-6, amsgrad=True, lr_step_size=16000, lr_gamma=0.5, model=m)
    x = torch.randn(Bsz, 4, S, S, S, device=dev)
    t = (torch.rand(Bsz, 3, S, S, S, device=dev) > 0.7).float()
    g = GraphedTrainStep(m, crit, opt, x, t, warmup=3)
    for _ in range(5):
        g()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3 if "--all" in sys.argv else 2):
            g()
        torch.cuda.synchronize()
path = out + ".trace.json"
prof.export_chrome_trace(path)
import json  # noqa: E402
ALL = "--all" in sys.argv
cats = ("kernel", "gpu_memset", "gpu_memcpy") if ALL else ("kernel",)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in cats]
ev.sort(key=lambda e: e["ts"])
if not ALL:
    # keep the second replay only
    half = len(ev) // 2
    ev = ev[half:]
t0 = ev[0]["ts"]
with open(out, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for e in ev:
        name = e["name"].split("(")[0].replace("void ", "").replace("b200::", "")
        f.write("%.3f,%.3f,%s,%s\n" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), name[:80]))
os.remove(path)
print("wrote", out, len(ev), "kernels, span %.1f us" % (ev[-1]["ts"] + ev[-1]["dur"] - t0))
