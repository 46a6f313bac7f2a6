"""Does processing level 0 ONE SAMPLE at a time keep its tensors in L2?  (perf diagnostic)

A 16-channel 128^3 bf16 tensor is 67 MB per sample: one sample fits the 126 MB L2, a batch of two (134 MB) does not.
Times a chain of K x [conv3 16->16 (+GN statistics) -> gn_finalize -> gn_apply(+lrelu)] on distinct buffers (as the
engine runs it: every stage writes a new tensor), for a batch of N = 2 in one go and for N = 1, and reports us per
sample per stage.  If N = 1 is clearly below half of N = 2, per-sample scheduling of level 0 would pay.

    python tools/l2_chain_probe.py [S] [K]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from brats2019_b200 import ops  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
Cc = 16
dev = "cuda"


def timed(fn, reps=10):
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


def chain(N):
    mk = lambda: ops.act_zeros(N, S, S, S, Cc, dev)
    acts = [mk() for _ in range(K + 1)]
    convs = [mk() for _ in range(K)]
    acts[0].interior().copy_(torch.randn(Cc // 8, N, S, S, S, 8, device=dev).to(torch.bfloat16))
    desc = ops.conv_desc(ops.MODE_K3, N, S, S, S, Cc, Cc)
    w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) / (Cc * 27) ** 0.5
    pk = ops.conv_pack_weight(desc, ops.W_FWD, w, K_real=Cc, N_real=Cc)
    ctas = ops.conv_ctas(desc)
    stats = torch.empty(ctas * N * 16, device=dev)
    mean = torch.empty(N * 8, device=dev); rstd = torch.empty(N * 8, device=dev)
    gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)

    def run():
        for k in range(K):
            ops.conv_run(desc, acts[k], pk, convs[k], stats=stats)
            ops.gn_finalize(stats, ctas, N, Cc, S, S, S, mean, rstd)
            ops.gn_apply(convs[k], mean, rstd, gamma, beta, acts[k + 1], lrelu=True)

    def run_conv():
        for k in range(K):
            ops.conv_run(desc, acts[k], pk, convs[k], stats=stats)

    def run_apply():
        for k in range(K):
            ops.gn_apply(convs[k], mean, rstd, gamma, beta, acts[k + 1], lrelu=True)

    with torch.cuda.stream(torch.cuda.Stream()):
        t, tc, ta = timed(run), timed(run_conv), timed(run_apply)
    torch.cuda.synchronize()
    return t / K / N, tc / K / N, ta / K / N


for N in (2, 1):
    t, tc, ta = chain(N)
    print("N=%d %d^3 C=%d: chain %.1f us per sample per stage | convs alone %.1f | gn_apply alone %.1f   (tensor %.0f MB per sample)"
          % (N, S, Cc, t, tc, ta, S ** 3 * Cc * 2 / 1e6))
