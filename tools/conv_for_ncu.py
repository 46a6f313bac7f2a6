"""The dominant kernel alone (3x3x3 C->C implicit GEMM at level `lvl` of a B x 4xS^3 step) for

    ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 3 -c 1 \
        -o gpurun_out/prof_conv python tools/conv_for_ncu.py [B] [S] [C]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
C = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
x = ops.act_zeros(B, S, S, S, C, dev)
x.interior().copy_(torch.randn(C // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
y = ops.act_zeros(B, S, S, S, C, dev)
w = torch.randn(C, C, 3, 3, 3, device=dev) * 0.05
desc = ops.conv_desc(ops.MODE_K3, B, S, S, S, C, C)
pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
st = torch.empty(ops.conv_ctas(desc) * B * 16, device=dev)
for _ in range(6):
    ops.conv_run(desc, x, pk, y, stats=st)
torch.cuda.synchronize()
