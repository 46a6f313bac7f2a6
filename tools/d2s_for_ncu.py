"""The k2 s2 data gradient with the depth-to-space epilogue alone (level l+1 -> l of a B x 4xS^3 step), for

    ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 3 -c 1 \
        -o gpurun_out/prof_d2s python tools/d2s_for_ncu.py [B] [S_coarse] [C_fine]
and, without ncu, its time alone (CUDA events, L2-cold: 3 x 134 MB of operands at the default shape).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 64          # coarse grid
Cf = int(sys.argv[3]) if len(sys.argv) > 3 else 16         # fine channels; coarse = 2 * Cf
dev = "cuda"
Cc = 2 * Cf
dy = ops.act_zeros(B, S, S, S, Cc, dev)
dy.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
skip = ops.act_zeros(B, 2 * S, 2 * S, 2 * S, Cf, dev)
skip.interior().copy_(torch.randn(Cf // 8, B, 2 * S, 2 * S, 2 * S, 8, device=dev).to(torch.bfloat16))
out = ops.act_zeros(B, 2 * S, 2 * S, 2 * S, Cf, dev)
w = torch.randn(Cc, Cf, 2, 2, 2, device=dev) * 0.05
desc = ops.conv_desc(ops.MODE_K1, B, S, S, S, Cc, 8 * Cf, epi=ops.EPI_D2S)
pk = ops.conv_pack_weight(desc, ops.W_DGRAD_S2D, w, K_real=Cc, N_real=8 * Cf)
for _ in range(6):
    ops.conv_run(desc, dy, pk, out, residual=skip)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    ops.conv_run(desc, dy, pk, out, residual=skip)
e1.record()
torch.cuda.synchronize()
nb = B * (2 * S) ** 3 * Cf * 2
print("d2s dgrad %d->8x%d @ %dx%d^3: %.1f us, %.0f GB/s algorithmic (dY + skip + out = %.0f MB)"
      % (Cc, Cf, B, S, e0.elapsed_time(e1) * 100, (2 * nb + nb // 4) / e0.elapsed_time(e1) * 10 / 1e6, (2 * nb + nb // 4) / 1e6))
