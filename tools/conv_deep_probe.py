"""Where does the time of the deep-level 3x3x3 convs go?  (probes build: skip-a-stage switches)

    python -m brats2019_b200.build --probes && python tools/conv_deep_probe.py

For 128 -> 128 @ 2x16^3 and 64 -> 64 @ 2x32^3 (the generic conv_gemm kernel): full kernel, no activation loads, no
MMAs, neither (= weight stream + epilogue).  Timed as 20 launches inside one CUDA graph (no launch gaps).
"""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ["B200_LIB_PATH"] = os.path.join(REPO, "brats2019_b200", "libbrats_b200_probes.so")
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
    for (B, s, Cc) in ((2, 16, 128), (2, 32, 64), (8, 16, 128)):
        x = ops.act_zeros(B, s, s, s, Cc, dev)
        x.interior().copy_(torch.randn(Cc // 8, B, s, s, s, 8, device=dev).to(torch.bfloat16))
        y = ops.act_zeros(B, s, s, s, Cc, dev)
        w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
        desc = ops.conv_desc(ops.MODE_K3, B, s, s, s, Cc, Cc)
        pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
        st = torch.empty(ops.conv_ctas(desc) * B * 16, device=dev)
        fl = 2.0 * B * s ** 3 * Cc * Cc * 27
        out = []
        for flag, what in ((0, "full"), (1, "no act loads"), (2, "no MMAs"), (3, "weights+epilogue only"), (3 + 16 + 32 + 64, "weights only")):
            os.environ["B200_CONV_DEBUG"] = str(flag)
            t = timed(lambda: ops.conv_run(desc, x, pk, y, stats=st))
            out.append("%s %.1f us" % (what, t))
        os.environ.pop("B200_CONV_DEBUG")
        print("conv3 %d->%d @ %dx%d^3 ctas=%d (%.1f GFLOP): %s" % (Cc, Cc, B, s, ops.conv_ctas(desc), fl / 1e9, " | ".join(out)), flush=True)
