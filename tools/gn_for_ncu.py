"""GroupNorm forward-apply / backward and the trilinear kernels alone at bench shapes, for ncu:
    ncu --set full --clock-control none --import-source on -k regex:'gn_|upsample' --launch-skip N -c M -o ... python tools/gn_for_ncu.py [B] [S] [C]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
x = ops.act_zeros(B, S, S, S, Cc, dev)
x.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
dy = ops.act_zeros(B, S, S, S, Cc, dev)
dy.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
y = ops.act_zeros(B, S, S, S, Cc, dev)
dx = ops.act_zeros(B, S, S, S, Cc, dev)
mean = torch.zeros(B * 8, device=dev); rstd = torch.ones(B * 8, device=dev)
gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
dg = torch.empty(Cc, device=dev); db = torch.empty(Cc, device=dev)
gws = ops.gn_backward_workspace(B, Cc, dev)
for _ in range(2):
    ops.gn_apply(x, mean, rstd, gamma, beta, y, residual=dy)
    ops.gn_backward(x, dy, mean, rstd, gamma, beta, dx, dg, db, gws)
    if S >= 32:
        co = ops.act_zeros(B, S // 2, S // 2, S // 2, Cc, dev)
        ops.upsample2x(co, y)
        ops.upsample2x_backward(y, y, co)
torch.cuda.synchronize()
