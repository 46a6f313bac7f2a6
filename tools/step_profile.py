"""Per-kernel device time of the steady-state training step (CUDA-graph replay) via torch.profiler / CUPTI:
    python tools/step_profile.py [B] [S] [steps]
Prints kernel families sorted by total time per step, the sum, and the graph's wall time per step (the
difference is idle time between kernels)."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import brats2019_b200 as B  # noqa: E402
from brats2019_b200.graphs import GraphedTrainStep  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        g()
    e1.record()
    torch.cuda.synchronize()
    wall = e0.elapsed_time(e1) / steps
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(steps):
            g()
        torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0].replace("void ", "").replace("b200::", "")
        if name.startswith("at::") or "multi_tensor" in name or "elementwise" in name:
            name = "torch: " + name[:60]
        tot[name][0] += 1
        tot[name][1] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
s = sum(v[1] for v in tot.values()) / steps
print("graph replay wall time %.3f ms/step; sum of kernel time %.3f ms/step (%d kernels/step)" % (
    wall, s / 1e3, sum(v[0] for v in tot.values()) // steps))
for name, (c, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:45]:
    print("%6.2f%%  %8.1f us  x%-4d %s" % (100 * us / steps / s, us / steps, c // steps, name[:100]))
