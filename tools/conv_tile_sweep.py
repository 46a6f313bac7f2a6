"""Tile-shape sweep of the generic implicit-GEMM conv at the deep levels (diagnostic): python tools/conv_tile_sweep.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

dev = "cuda"
for (B, S, Cc) in ((2, 16, 128), (2, 32, 64)):
    x = ops.act_zeros(B, S, S, S, Cc, dev)
    x.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
    w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
    out = ops.act_zeros(B, S, S, S, Cc, dev)
    flops = 2.0 * B * S ** 3 * Cc * Cc * 27
    for tile in ("", "1 1 0", "1 2 0", "2 1 0", "1 1 1", "1 2 1", "1 4 0", "4 1 0", "2 2 0"):
        if tile:
            os.environ["B200_CONV_TILE"] = tile
        else:
            os.environ.pop("B200_CONV_TILE", None)
        try:
            desc = ops.conv_desc(ops.MODE_K3, B, S, S, S, Cc, Cc)
            pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
            st = torch.empty(ops.conv_ctas(desc) * B * 16, device=dev)
            for _ in range(3):
                ops.conv_run(desc, x, pk, out, stats=st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                ops.conv_run(desc, x, pk, out, stats=st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            print("conv3 %d->%d @ %dx%d^3 tile[BD MB whole]=%-8s ctas %3d  %.4f ms  %.0f TFLOP/s" % (
                Cc, Cc, B, S, tile or "planner", ops.conv_ctas(desc), ms, flops / ms / 1e9))
        except RuntimeError as e:
            print("tile", tile, "->", str(e)[:80])
