"""Does a weight-gradient GEMM overlap with the memory-bound kernels when queued on another stream?
(perf diagnostic, GPU box):  python tools/overlap_probe.py [B] [S] [C]"""
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
x, dy, dx, x2, dy2 = mk(), mk(), mk(), mk(), mk()
for a in (x, dy, x2, dy2):
    a.interior().copy_(torch.randn(Cc // 8, B, S, S, S, 8, device=dev).to(torch.bfloat16))
mean = torch.zeros(B * 8, device=dev); rstd = torch.ones(B * 8, device=dev)
gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
dg = torch.empty(Cc, device=dev); db = torch.empty(Cc, device=dev)
gws = ops.gn_backward_workspace(B, Cc, dev)
g = torch.empty(Cc, Cc, 3, 3, 3, device=dev)
wd = ops.wgrad_desc(0, B, S, S, S, Cc, Cc)
ws = ops.wgrad_workspace(wd, dev)
w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
cd = ops.conv_desc(ops.MODE_K3, B, S, S, S, Cc, Cc)
pk = ops.conv_pack_weight(cd, ops.W_FWD, w)
y = mk()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def gnb():
    ops.gn_backward(x, dy, mean, rstd, gamma, beta, dx, dg, db, gws)


def wg():
    ops.wgrad_run(wd, dy2, x2, g, ops.G_K3, workspace=ws)


def conv():
    ops.conv_run(cd, x2, pk, y)


def timed(fa, fb, reps=20):
    """fa on stream 1 and (optionally) fb on stream 2, inside one CUDA graph so that launch overhead is out."""
    gr = torch.cuda.CUDAGraph()
    for _ in range(2):
        fa()
        if fb:
            fb()
    torch.cuda.synchronize()
    with torch.cuda.graph(gr):
        cur = torch.cuda.current_stream()
        for _ in range(reps):
            if fb:
                s2.wait_stream(cur)
                with torch.cuda.stream(s2):
                    fb()
            fa()
            if fb:
                cur.wait_stream(s2)
    gr.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    gr.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


big_a = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
big_b = torch.empty_like(big_a)


def copy():
    big_b.copy_(big_a)


def gnf():
    ops.gn_apply(x, mean, rstd, gamma, beta, dx, residual=dy)


with torch.cuda.stream(s1):
    t_copy, t_gnf = timed(copy, None), timed(gnf, None)
    t_wg = timed(wg, None)
    print("copy 512MB %.1f us | gn_apply %.1f us | wgrad %.1f us | copy || wgrad %.1f us | gn_apply || wgrad %.1f us | wgrad || wgrad %.1f us"
          % (t_copy, t_gnf, t_wg, timed(copy, wg), timed(gnf, wg), timed(wg, wg)))
    a = timed(gnb, None)
    b = timed(wg, None)
    c = timed(conv, None)
    ab = timed(gnb, wg)
    cb = timed(conv, wg)
print("C=%d %dx%d^3: gn_backward %.1f us | wgrad %.1f us | conv %.1f us | gn_backward || wgrad %.1f us (sum %.1f) | conv || wgrad %.1f us (sum %.1f)"
      % (Cc, B, S, a, b, c, ab, a + b, cb, c + b))
