"""Does a consumer kernel hit in L2 on what the previous kernel wrote?  (hardware micro-test, torch ops only)"""
import torch
dev = "cuda"
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def t(fn, pre, reps=20):
    ts = []
    for _ in range(reps):
        pre()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for mb in (8, 28, 48, 67, 100, 134):
    n = mb * (1 << 20) // 2
    src = torch.randn(n, device=dev).to(torch.bfloat16)
    a = torch.empty_like(src); out = torch.empty_like(src)
    read = lambda: torch.add(a, 1.0, out=out)          # reads a (just written or cold), writes out
    cold = t(read, lambda: (a.copy_(src), flush.zero_()))
    warm_w = t(read, lambda: a.copy_(src))             # a was just WRITTEN by the previous kernel
    warm_r = t(read, lambda: (a.copy_(src), flush.zero_(), a.sum()))   # a was just READ by the previous kernel
    print("%4d MB: read+write pass cold %.1f us (%.0f GB/s) | after a write of it %.1f us | after a read of it %.1f us"
          % (mb, cold, 2 * n * 2 / cold / 1e3, warm_w, warm_r))
