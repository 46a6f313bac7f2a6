"""Kernel-level timing table (diagnostic, run on the GPU box): every kernel family at the shapes
of the benchmark (batch B x 4x128^3), CUDA events, 3 warm-ups, L2-cold inputs are not needed for
the large levels (tensors >> L2) and noted for the small ones.

    python tools/perf_probe.py [B] [S]
Prints `name  ms  TFLOP/s | GB/s  (fraction of measured peak)`.
"""
import json
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
# the skip-a-stage / cycle-accounting code exists only in the probes build: python -m brats2019_b200.build --probes
_PROBES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "brats2019_b200", "libbrats_b200_probes.so")
if "--probes" in sys.argv:
    if not os.path.exists(_PROBES):
        raise SystemExit("build the probes library first: python -m brats2019_b200.build --probes")
    os.environ["B200_LIB_PATH"] = _PROBES


import torch  # noqa: E402


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops"]
    return 6650.0, 1590.0


class Ms(float):
    """Back-to-back time per launch; `.cold` = one launch at a time after an L2 flush (a 256 MB write), for the
    memory-bound rows: a probe that re-runs on the same buffers is partly L2-fed (126 MB L2) and can read as > 1.0 of
    the HBM peak - the cold number is the one compared with the roofline."""
    cold = None


_FLUSH = None


def timeit(fn, reps=10, warm=3, cold_reps=5):
    global _FLUSH
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = Ms(e0.elapsed_time(e1) / reps)
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    tot = 0.0
    for _ in range(cold_reps):
        _FLUSH.zero_()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms.cold = tot / cold_reps
    return ms


def line(name, ms, flops=None, bytes_=None):
    hbm, tf = peaks()
    s = "%-46s %8.3f ms" % (name, ms)
    if flops:
        t = flops / ms / 1e9
        s += "  %8.1f TFLOP/s (%.3f of %.0f)" % (t, t / tf, tf)
    if bytes_:
        cold = getattr(ms, "cold", None)
        if cold:
            g = bytes_ / cold / 1e6
            s += "  L2-cold %6.3f ms %8.1f GB/s (%.3f of %.0f)" % (cold, g, g / hbm, hbm)
        else:
            g = bytes_ / ms / 1e6
            s += "  %8.1f GB/s (%.3f of %.0f)" % (g, g / hbm, hbm)
    print(s, flush=True)


def main():
    from brats2019_b200 import ops
    B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 128
    dev = "cuda"
    ch = [16, 32, 64, 128]
    only = os.environ.get("PROBE_ONLY", "")
    for lvl, Cc in enumerate(ch):
        s = S >> lvl
        vox = B * s ** 3
        x = ops.act_zeros(B, s, s, s, Cc, dev)
        x.interior().copy_(torch.randn(Cc // 8, B, s, s, s, 8, device=dev).to(torch.bfloat16))
        y = ops.act_zeros(B, s, s, s, Cc, dev)
        w = torch.randn(Cc, Cc, 3, 3, 3, device=dev) * 0.05
        desc = ops.conv_desc(ops.MODE_K3, B, s, s, s, Cc, Cc)
        pk = ops.conv_pack_weight(desc, ops.W_FWD, w)
        st = torch.empty(ops.conv_ctas(desc) * B * 16, device=dev)
        fl = 2.0 * vox * Cc * Cc * 27
        if not only or "conv3" in only:
            line("conv3 %d->%d @ %dx%d^3 (+GN stats)" % (Cc, Cc, B, s), timeit(lambda: ops.conv_run(desc, x, pk, y, stats=st)), flops=fl)
        if lvl == 0 and (not only or "conv3" in only) and os.environ.get("B200_CONV_DEBUG") is None and "--probes" in sys.argv:
            for flag, what in ((1, "no activation loads"), (2, "no MMAs"), (3, "neither (epilogue + weights only)"),
                               (7, "neither, no fold exchange/barrier"), (15, "neither, no exchange, no shuffles"),
                               (3 + 16, "neither, no tmem ld"), (3 + 32, "neither, no stores"), (3 + 64, "neither, no stats"),
                               (15 + 16, "neither, no xch/shfl/tmem ld"), (15 + 32, "neither, no xch/shfl/stores"),
                               (15 + 16 + 32 + 64, "epilogue skeleton only"), (16 + 32 + 64 + 12, "loads+MMA, skeleton epilogue")):
                env = dict(os.environ, B200_CONV_DEBUG=str(flag), PROBE_ONLY="conv3L0")
                r = subprocess.run([sys.executable, os.path.abspath(__file__), str(B), str(S)], env=env, capture_output=True, text=True)
                print("   probe[%s]: %s" % (what, r.stdout.strip().splitlines()[0] if r.stdout.strip() else r.stderr[-300:]), flush=True)
        if only == "conv3L0":
            return
        if only:
            continue
        # wgrad
        dy = ops.act_zeros(B, s, s, s, Cc, dev)
        dy.interior().copy_(torch.randn(Cc // 8, B, s, s, s, 8, device=dev).to(torch.bfloat16))
        g = torch.empty(Cc, Cc, 3, 3, 3, device=dev)
        wd = ops.wgrad_desc(0, B, s, s, s, Cc, Cc)
        ws = ops.wgrad_workspace(wd, dev)
        line("wgrad3 %dx%d @ %dx%d^3" % (Cc, Cc, B, s), timeit(lambda: ops.wgrad_run(wd, dy, x, g, ops.G_K3, workspace=ws)), flops=fl)
        # GN apply / backward
        mean = torch.zeros(B * 8, device=dev); rstd = torch.ones(B * 8, device=dev)
        gamma = torch.ones(Cc, device=dev); beta = torch.zeros(Cc, device=dev)
        nb = vox * Cc * 2
        line("gn_apply+lrelu+residual C=%d" % Cc, timeit(lambda: ops.gn_apply(x, mean, rstd, gamma, beta, y, residual=dy)), bytes_=3 * nb)   # x + residual read, y written
        dx = ops.act_zeros(B, s, s, s, Cc, dev)
        dg = torch.empty(Cc, device=dev); db = torch.empty(Cc, device=dev)
        gws = ops.gn_backward_workspace(B, Cc, dev)
        line("gn_backward (3 kernels) C=%d" % Cc, timeit(lambda: ops.gn_backward(x, dy, mean, rstd, gamma, beta, dx, dg, db, gws)), bytes_=5 * nb)
        if lvl > 0:
            fine = ops.act_zeros(B, 2 * s, 2 * s, 2 * s, Cc, dev)
            line("upsample2x+lrelu C=%d -> %d^3" % (Cc, 2 * s), timeit(lambda: ops.upsample2x(x, fine)), bytes_=9 * nb)
            dco = ops.act_zeros(B, s, s, s, Cc, dev)
            line("upsample2x backward C=%d" % Cc, timeit(lambda: ops.upsample2x_backward(fine, fine, dco)), bytes_=17 * nb)
        if lvl < 3:
            co = ops.act_zeros(B, s // 2, s // 2, s // 2, 8 * Cc, dev)
            line("space_to_depth C=%d" % Cc, timeit(lambda: ops.space_to_depth(x, co)), bytes_=2 * nb)
            wdn = torch.randn(2 * Cc, Cc, 2, 2, 2, device=dev) * 0.05
            d1 = ops.conv_desc(ops.MODE_K1, B, s // 2, s // 2, s // 2, 8 * Cc, 2 * Cc)
            pk1 = ops.conv_pack_weight(d1, ops.W_FWD_S2D, wdn)
            o1 = ops.act_zeros(B, s // 2, s // 2, s // 2, 2 * Cc, dev)
            line("conv1 (down) %d->%d @ %d^3" % (8 * Cc, 2 * Cc, s // 2), timeit(lambda: ops.conv_run(d1, co, pk1, o1)),
                 flops=2.0 * vox / 8 * 8 * Cc * 2 * Cc, bytes_=(vox / 8) * (8 * Cc + 2 * Cc) * 2)
            wcat = torch.randn(Cc, 2 * Cc, 1, 1, 1, device=dev) * 0.05
            d2 = ops.conv_desc(ops.MODE_K1, B, s, s, s, Cc, Cc, Cin_b=Cc)
            pk2 = ops.conv_pack_weight(d2, ops.W_FWD, wcat)
            line("conv1 cat %d+%d->%d @ %d^3" % (Cc, Cc, Cc, s), timeit(lambda: ops.conv_run(d2, x, pk2, y, src_b=dy)),
                 flops=2.0 * vox * 2 * Cc * Cc, bytes_=3 * nb)
    if only:
        return
    # input packing, sigmoid head, dice
    xin = torch.randn(B, 4, S, S, S, device=dev)
    x16 = ops.act_zeros(B, S, S, S, 16, dev)
    line("pack_input 4->16", timeit(lambda: ops.pack_input(xin, 16, out=x16)), bytes_=B * S ** 3 * (16 + 32))
    probs = torch.rand(B, 3, S, S, S, device=dev); tgt = (torch.rand(B, 3, S, S, S, device=dev) > 0.7).float()
    sums = torch.zeros(8, device=dev)
    line("dice sums", timeit(lambda: ops.dice_sums(probs, tgt, sums)), bytes_=B * 3 * S ** 3 * 8)
    gp = torch.empty_like(probs); one = torch.ones(1, device=dev)
    line("dice backward", timeit(lambda: ops.dice_backward(probs, tgt, sums, one, 1.0, gp)), bytes_=B * 3 * S ** 3 * 12)
    dl = ops.act_zeros(B, S, S, S, 16, dev); dbias = torch.empty(3, device=dev)
    line("sigmoid backward -> act16", timeit(lambda: ops.sigmoid_backward(gp, probs, dl, dbias)), bytes_=B * S ** 3 * (24 + 32))


if __name__ == "__main__":
    main()
