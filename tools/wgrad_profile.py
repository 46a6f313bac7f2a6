"""Per-kernel device time of one weight-gradient call (gemm + reduce) via torch.profiler:
    python tools/wgrad_profile.py [B] [S] [C]"""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
x = ops.act_zeros(B, S, S, S, Cc, dev)
dy = ops.act_zeros(B, S, S, S, Cc, dev)
for a in (x, dy):
    a.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
g = torch.empty(Cc, Cc, 3, 3, 3, device=dev)
wd = ops.wgrad_desc(0, B, S, S, S, Cc, Cc)
ws = ops.wgrad_workspace(wd, dev)
for _ in range(5):
    ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws)
torch.cuda.synchronize()
reps = 10
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(reps):
        ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws)
    torch.cuda.synchronize()
tot = defaultdict(float)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        tot[ev.name.split("(")[0]] += ev.device_time
print("C=%d %dx%d^3: " % (Cc, B, S) + " | ".join("%s %.1f us" % (k.replace("b200::", ""), v / reps) for k, v in tot.items()))
