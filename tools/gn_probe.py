"""GroupNorm kernel timings at one level (perf diagnostic): python tools/gn_probe.py [B] [S] [C]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
Cc = int(sys.argv[3]) if len(sys.argv) > 3 else 16
dev = "cuda"
mk = lambda: ops.act_zeros(B, S, S, S, Cc, dev)
x, dy, dx, y = mk(), mk(), mk(), mk()
for a in (x, dy):
    a.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
mean = torch.zeros(B * 8, device=dev); rstd = torch.ones(B * 8, device=dev)
gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
dg = torch.empty(Cc, device=dev); db = torch.empty(Cc, device=dev)
gws = ops.gn_backward_workspace(B, Cc, dev)


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


nb = B * S ** 3 * Cc * 2
with torch.cuda.stream(torch.cuda.Stream()):
    t1 = timed(lambda: ops.gn_apply(x, mean, rstd, gamma, beta, y, lrelu=True))
    t2 = timed(lambda: ops.gn_apply(x, mean, rstd, gamma, beta, y, residual=dy, lrelu=True))
    t3 = timed(lambda: ops.gn_backward(x, dy, mean, rstd, gamma, beta, dx, dg, db, gws))
print("C=%d %dx%d^3 (tensor %.1f MB): gn_apply %.1f us (%.0f GB/s) | gn_apply+res %.1f us (%.0f GB/s) | gn_backward %.1f us (%.0f GB/s)"
      % (Cc, B, S, nb / 1e6, t1, 2 * nb / t1 / 1e3, t2, 3 * nb / t2 / 1e3, t3, 5 * nb / t3 / 1e3))
