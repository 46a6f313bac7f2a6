"""Where the line-marching weight-gradient kernel spends its time (perf diagnostic, GPU box, probes build):
    python -m brats2019_b200.build --probes && python tools/wgl_prof.py [B] [S] [Cout_w] [Cin_w]
Times the kernel with one stage switched off at a time (B200_WGL_DEBUG bits: 1 no X copies, 2 no MMAs, 4 no shift
copies, 8 no dY copies) and prints the MMA warp's cycle accounting (bit 256)."""
import ctypes as C
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
os.environ["B200_LIB_PATH"] = os.path.join(REPO, "brats2019_b200", "libbrats_b200_probes.so")
import torch  # noqa: E402

from brats2019_b200 import _lib, ops  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    dev = "cuda"
    x = ops.act_zeros(B, S, S, S, 16, dev)
    dy = ops.act_zeros(B, S, S, S, 16, dev)
    x.interior().copy_(torch.randn(2, B, S, S, S, 8, device=dev).to(torch.bfloat16))
    dy.interior().copy_(torch.randn(2, B, S, S, S, 8, device=dev).to(torch.bfloat16))
    desc = ops.wgrad_desc(0, B, S, S, S, 16, 16)
    ws = ops.wgrad_workspace(desc, dev)
    co_w = int(sys.argv[3]) if len(sys.argv) > 3 else 16       # PyTorch channel counts of the gradient: <= 8 on one side
    ci_w = int(sys.argv[4]) if len(sys.argv) > 4 else 16       # skips that side's upper chunk (conv_input: 16 4, conv_output: 3 16)
    g = torch.zeros(co_w, ci_w, 3, 3, 3, device=dev)
    print("gradient %d x %d" % (co_w, ci_w))
    L = _lib.lib()
    L.b200_wgl_prof_read.restype = C.c_int
    L.b200_wgl_prof_read.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    for dbg, what in ((0, "full kernel"), (2, "no MMAs"), (1, "no X copies"), (8, "no dY copies"), (4, "no shift copies"),
                      (9, "no copies from L2 at all"), (13, "MMAs only (no copies, no shifts)"), (15, "skeleton (barriers only)")):
        os.environ["B200_WGL_DEBUG"] = str(dbg | 256)
        for _ in range(2):
            ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            ops.wgrad_run(desc, dy, x, g, ops.G_K3, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        buf = (C.c_ulonglong * (160 * 8))()
        _lib.check(L.b200_wgl_prof_read(buf, 160 * 8), "prof_read")
        n = 148
        tot = sum(buf[c * 8 + 0] for c in range(n)) / n
        wx = sum(buf[c * 8 + 1] for c in range(n)) / n
        wy = sum(buf[c * 8 + 2] for c in range(n)) / n
        ti = sum(buf[c * 8 + 3] for c in range(n)) / n
        steps = sum(buf[c * 8 + 4] for c in range(n)) / n
        print("%-36s %.4f ms | MMA warp: total %7.0f cycles = wait X %7.0f + wait dY %7.0f + issue %7.0f (%.0f steps, %.0f cycles/step)"
              % (what, e0.elapsed_time(e1) / 10, tot, wx, wy, ti, steps, tot / max(steps, 1)))


if __name__ == "__main__":
    main()
