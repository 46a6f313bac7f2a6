"""A small training step + the op checks that touch this round's kernel changes, for compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import brats2019_b200 as B  # noqa: E402

torch.manual_seed(0)
m = B.UNet(**B.DEFAULT_CFG).cuda().train()
crit = B.Dice_loss_joint()
shape = tuple(int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (1, 16, 32, 32)
x = torch.randn(shape[0], 4, *shape[1:], device="cuda")
t = (torch.rand(shape[0], 3, *shape[1:], device="cuda") > 0.7).float()
for env in ({}, {"B200_D2S_V8": "1", "B200_WGL_PAIR": "7"}):
    os.environ.update(env)
    for _ in range(2):
        for p in m.parameters():
            p.grad = None
        loss = crit(m([x]), [t])
        loss.backward()
    torch.cuda.synchronize()
    print("step ok", env, float(loss))
