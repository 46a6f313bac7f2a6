"""A/B timing of the 16 x 16 weight-gradient kernels alone (CUDA events, L2 flushed by size: 2 x 134 MB operands):
    python tools/wgrad_ab.py [B] [S]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 128
    dev = "cuda"
    x = ops.act_zeros(B, S, S, S, 16, dev)
    dy = ops.act_zeros(B, S, S, S, 16, dev)
    x.interior().copy_(torch.randn(2, B, S, S, S, 8, device=dev).to(torch.bfloat16))
    dy.interior().copy_(torch.randn(2, B, S, S, S, 8, device=dev).to(torch.bfloat16))
    desc = ops.wgrad_desc(0, B, S, S, S, 16, 16)
    ws = ops.wgrad_workspace(desc, dev)
    flops = 2.0 * B * S ** 3 * 27 * 16 * 16
    res = {}
    sweeps = [("line", {"B200_NO_WGRAD_LINE": "0", "B200_WGL_PAIR": "0"}), ("linear", {"B200_NO_WGRAD_LINE": "1"}),
              ("line pair mode", {"B200_NO_WGRAD_LINE": "0", "B200_WGL_PAIR": "7"})]
    if "--pair-sweep" in sys.argv:
        sweeps += [("pair LH=%d Ny=%d" % (lh, ny), {"B200_NO_WGRAD_LINE": "0", "B200_WGL_PAIR": "7", "B200_WGL_LH": str(lh),
                                                   "B200_WGL_NR": "0", "B200_WGL_NY": str(ny)})
                   for lh, ny in ((4, 6), (4, 4), (6, 4), (2, 8), (4, 8))]
    if "--sweep" in sys.argv:
        sweeps += [("line LH=%d NR=%d Ny=%d" % (lh, nr, ny), {"B200_NO_WGRAD_LINE": "0", "B200_WGL_LH": str(lh), "B200_WGL_NR": str(nr),
                                                            "B200_WGL_NY": str(ny)})
                   for lh, nr, ny in ((8, 8, 3), (8, 6, 3), (8, 4, 3), (8, 3, 3), (8, 2, 2), (8, 4, 2), (8, 8, 2), (6, 8, 4), (6, 4, 3))]
    for form, env in sweeps:
        for k in ("B200_WGL_LH", "B200_WGL_NY", "B200_WGL_NR"):
            os.environ.pop(k, None)
        os.environ.update(env)
        g = torch.zeros(16, 16, 3, 3, 3, device=dev)
        for _ in range(3):
            ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(20):
            ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        res[form] = g.clone()
        print("wgrad3 16x16 @ %dx%d^3 %-22s %.4f ms (gemm + reduce)  %.1f TFLOP/s" % (B, S, form, ms, flops / ms / 1e9))
    if "--contend" in sys.argv:
        # the same two kernels while a memory-bound kernel (a 1 GiB device copy per launch) runs on another stream, as the
        # GroupNorm-backward kernels do beside the weight gradients in the training step
        a = torch.empty(1 << 28, dtype=torch.float32, device=dev)
        b = torch.empty_like(a)
        s2 = torch.cuda.Stream()
        for form, env in sweeps[:2]:
            os.environ.pop("B200_WGL_LH", None); os.environ.pop("B200_WGL_NY", None)
            os.environ.update(env)
            g = torch.zeros(16, 16, 3, 3, 3, device=dev)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(s2):
                c0.record()
                for _ in range(6):
                    b.copy_(a)
                c1.record()
            e0.record()
            for _ in range(20):
                ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
            e1.record()
            torch.cuda.synchronize()
            print("beside a device copy: %-8s %.4f ms per wgrad (20 launches), the 6 GiB of copies took %.3f ms (alone: ~%.3f ms)"
                  % (form, e0.elapsed_time(e1) / 20, c0.elapsed_time(c1), 6 * 2 * (1 << 30) / 6.5e9))
    for k in ("B200_WGL_LH", "B200_WGL_NY", "B200_WGL_NR"):
        os.environ.pop(k, None)
    for name, r in res.items():
        if name != "linear":
            print("%s vs linear rel-L2 %.2e" % (name, ((r - res["linear"]).norm() / res["linear"].norm()).item()))
    # the 4-input-channel (conv_input) and 3-output-channel (conv_output) variants
    for (co, ci) in ((16, 4), (3, 16)):
        for form, env in sweeps[:3]:
            os.environ.update(env)
            g = torch.zeros(co, ci, 3, 3, 3, device=dev)
            for _ in range(3):
                ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(20):
                ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
            e1.record()
            torch.cuda.synchronize()
            print("wgrad3 %dx%d %-22s %.4f ms" % (co, ci, form, e0.elapsed_time(e1) / 20))


if __name__ == "__main__":
    main()
