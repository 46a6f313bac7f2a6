"""One training step (or one forward) of the drop-in UNet between cudaProfilerStart/Stop, for

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/step_for_ncu.py [B] [S] [fwd]

(launch list of exactly one step; kernels are launched one by one, no CUDA graph)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brats2019_b200 as B  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
fwd_only = len(sys.argv) > 3 and sys.argv[3] == "fwd"
torch.manual_seed(0)
m = B.UNet(**B.DEFAULT_CFG).cuda().train()
crit = B.Dice_loss_joint()
from brats2019_b200.optim import FusedAdam  # noqa: E402
opt = FusedAdam(m.parameters(), lr=2e-5, weight_decay=1e-6, amsgrad=True, lr_step_size=16000, lr_gamma=0.5, model=m)
x = torch.randn(Bsz, 4, S, S, S, device="cuda")
t = (torch.rand(Bsz, 3, S, S, S, device="cuda") > 0.7).float()


def step():
    if fwd_only:
        with torch.no_grad():
            m([x])
        return
    opt.zero_grad(set_to_none=True)
    loss = crit(m([x]), [t])
    loss.backward()
    opt.step()


for _ in range(2):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
