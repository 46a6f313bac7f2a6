"""Op-level GPU diagnostics: every CUDA kernel family against a plain torch fp32 reference.

    python tests/gpu_opcheck.py [group ...]        # groups: ew conv3 conv1 dgrad wgrad

Each group runs in its own subprocess (a device fault in one group must not hide the others)
and prints one line per check: `PASS|FAIL name  max_abs_err  ref_max  details`.
`tests/test_ops_gpu.py` wraps the same checks for pytest.  The torch ops used here are the
checker only (the "plain PyTorch fp32 reference of the same op"), never the product path.
"""
import os
import subprocess
import sys
import traceback

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

RESULTS = []


def bf(x):
    return x.to(__import__("torch").bfloat16).float()


def report(name, got, ref, tol_rel=2e-2, tol_abs=None, extra=""):
    import torch
    got = got.float()
    ref = ref.float()
    refmax = ref.abs().max().item()
    err = (got - ref).abs()
    maxerr = err.max().item()
    tol = tol_rel * max(refmax, 1e-6) if tol_abs is None else tol_abs
    bad = err > tol
    ok = (not bad.any().item()) and bool(torch.isfinite(got).all().item())
    line = "%s %-44s err=%.4g ref_max=%.4g tol=%.3g bad=%d/%d %s" % (
        "PASS" if ok else "FAIL", name, maxerr, refmax, tol, int(bad.sum().item()), bad.numel(), extra)
    if not ok:
        idx = torch.nonzero(bad)[:6].tolist()
        samples = ["%s got=%.4g ref=%.4g" % (tuple(i), got[tuple(i)].item(), ref[tuple(i)].item()) for i in idx]
        line += "\n      nonfinite=%d zero_out_frac=%.3f first_bad: %s" % (
            int((~torch.isfinite(got)).sum().item()), float((got == 0).float().mean().item()), "; ".join(samples))
    print(line, flush=True)
    RESULTS.append(ok)
    return ok


# ----------------------------------------------------------------------------------------------
def group_ew():
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(0)
    N, D, H, W = 2, 4, 6, 10
    # pack_input
    x = torch.randn(N, 4, D, H, W, device=dev)
    a = ops.pack_input(x, 16)
    report("pack_input interior", ops.act_to_ncdhw(a, 4), bf(x), tol_abs=0)
    report("pack_input halo+guard zero", ops.act_outside_absmax(a).view(1), torch.zeros(1, device=dev), tol_abs=0)
    report("pack_input pad channels zero", ops.act_to_ncdhw(a)[:, 4:], torch.zeros(N, 12, D, H, W, device=dev), tol_abs=0)
    # W % 4 == 0 takes the four-voxels-per-thread kernel (W = 10 above takes the scalar one); odd line count per CTA
    x4 = torch.randn(3, 4, 5, 7, 24, device=dev)
    a4 = ops.pack_input(x4, 16)
    report("pack_input quad interior", ops.act_to_ncdhw(a4, 4), bf(x4), tol_abs=0)
    report("pack_input quad halo+guard zero", ops.act_outside_absmax(a4).view(1), torch.zeros(1, device=dev), tol_abs=0)
    report("pack_input quad pad channels zero", ops.act_to_ncdhw(a4)[:, 4:], torch.zeros(3, 12, 5, 7, 24, device=dev), tol_abs=0)

    for Cc in (16, 32, 64, 128):
        xc = bf(torch.randn(N, Cc, D, H, W, device=dev) * 2 + 0.5)
        xa = ops.act_from_ncdhw(xc)
        gamma = torch.randn(Cc, device=dev) * 0.5 + 1
        beta = torch.randn(Cc, device=dev) * 0.3
        res = bf(torch.randn(N, Cc, D, H, W, device=dev))
        ra = ops.act_from_ncdhw(res)
        # stats as the conv epilogue would produce them: [ctas][N][16]
        g = xc.view(N, 8, -1)
        stats = torch.zeros(3, N, 16, device=dev)
        stats[1, :, :8] = g.sum(-1)
        stats[1, :, 8:] = (g * g).sum(-1)
        mean = torch.empty(N * 8, device=dev); rstd = torch.empty(N * 8, device=dev)
        ops.gn_finalize(stats, 3, N, Cc, D, H, W, mean, rstd)
        report("gn_finalize mean C=%d" % Cc, mean.view(N, 8), g.mean(-1), tol_abs=1e-4)
        report("gn_finalize rstd C=%d" % Cc, rstd.view(N, 8), 1 / torch.sqrt(g.var(-1, unbiased=False) + 1e-5), tol_rel=1e-4)
        out = ops.act_zeros(N, D, H, W, Cc, dev)
        ops.gn_apply(xa, mean, rstd, gamma, beta, out, residual=ra, lrelu=True)
        ref = F.leaky_relu(F.group_norm(xc, 8, gamma, beta, 1e-5), 0.01) + res
        report("gn_apply+lrelu+res C=%d" % Cc, ops.act_to_ncdhw(out), ref, tol_rel=1e-2)
        ops.gn_apply(xa, mean, rstd, gamma, beta, out, residual=None, lrelu=False)
        report("gn_apply plain C=%d" % Cc, ops.act_to_ncdhw(out), F.group_norm(xc, 8, gamma, beta, 1e-5), tol_rel=1e-2)
        # backward
        for lre in (True, False):
            xr = xc.clone().requires_grad_(True)
            gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
            y = F.group_norm(xr, 8, gr, br, 1e-5)
            if lre:
                y = F.leaky_relu(y, 0.01)
            dy = bf(torch.randn_like(y))
            y.backward(dy)
            # both forms behind b200_gn_backward: the one-launch cluster kernel (small tensors) and the
            # reduce -> finalize -> apply path (B200_GN_BWD_CLUSTER=0 forces it)
            for form in ("creg", "cluster", "3-kernel"):
                # register-resident cluster kernel (default for small tensors), shared-memory cluster kernel, three launches
                os.environ["B200_GN_BWD_CREG"] = "1" if form == "creg" else "0"
                os.environ["B200_GN_BWD_CLUSTER"] = "1" if form == "cluster" else "0"
                dx = ops.act_zeros(N, D, H, W, Cc, dev)
                dgam = torch.empty(Cc, device=dev); dbet = torch.empty(Cc, device=dev)
                ws = ops.gn_backward_workspace(N, Cc, dev)
                ops.gn_backward(xa, ops.act_from_ncdhw(dy), mean, rstd, gamma, beta, dx, dgam, dbet, ws, lrelu=lre)
                tag = "C=%d lrelu=%d %s" % (Cc, lre, form)
                report("gn_backward dx " + tag, ops.act_to_ncdhw(dx), xr.grad, tol_rel=2e-2)
                report("gn_backward dgamma " + tag, dgam, gr.grad, tol_rel=2e-3)
                report("gn_backward dbeta " + tag, dbet, br.grad, tol_rel=2e-3)
                report("gn_backward halo untouched " + tag, ops.act_outside_absmax(dx).view(1), torch.zeros(1, device=dev), tol_abs=0)
            os.environ.pop("B200_GN_BWD_CLUSTER", None)
            os.environ.pop("B200_GN_BWD_CREG", None)
        # many reduction CTAs per sample, odd batch: the last-CTA finalize inside the reduction kernel (default) against
        # the separate finalize launch (B200_GN_BWD_FUSED_FIN=0) and against autograd; repeated calls reuse the tickets
        for (N2, D2, H2, W2) in (((3, 32, 32, 32) if Cc == 16 else (2, 16, 16, 16)),):
            x2 = bf(torch.randn(N2, Cc, D2, H2, W2, device=dev) * 2 + 0.5)
            g2 = x2.view(N2, 8, -1)
            mean2 = g2.mean(-1).reshape(-1).contiguous()
            rstd2 = (1 / torch.sqrt(g2.var(-1, unbiased=False) + 1e-5)).reshape(-1).contiguous()
            xr = x2.clone().requires_grad_(True)
            gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
            y = F.leaky_relu(F.group_norm(xr, 8, gr, br, 1e-5), 0.01)
            dy2 = bf(torch.randn_like(y))
            y.backward(dy2)
            xa2, dya2 = ops.act_from_ncdhw(x2), ops.act_from_ncdhw(dy2)
            ws = ops.gn_backward_workspace(N2, Cc, dev)
            res = {}
            os.environ["B200_GN_BWD_CREG"] = "0"         # (these shapes would otherwise take the one-launch cluster kernel)
            for form in ("fused", "fused again", "separate"):
                os.environ["B200_GN_BWD_FUSED_FIN"] = "0" if form == "separate" else "1"
                dx = ops.act_zeros(N2, D2, H2, W2, Cc, dev)
                dgam = torch.full((Cc,), 5.0, device=dev); dbet = torch.full((Cc,), 5.0, device=dev)
                ops.gn_backward(xa2, dya2, mean2, rstd2, gamma, beta, dx, dgam, dbet, ws, lrelu=True)
                torch.cuda.synchronize()
                res[form] = (ops.act_to_ncdhw(dx), dgam, dbet)
                tag = "C=%d %dx(%d,%d,%d) %s" % (Cc, N2, D2, H2, W2, form)
                report("gn_backward dx " + tag, res[form][0], xr.grad, tol_rel=2e-2)
                report("gn_backward dgamma " + tag, dgam, gr.grad, tol_rel=2e-3)
                report("gn_backward dbeta " + tag, dbet, br.grad, tol_rel=2e-3)
            os.environ.pop("B200_GN_BWD_FUSED_FIN", None)
            os.environ["B200_GN_BWD_CREG"] = "1"
            # ... and the register-resident one-launch kernel at the same shape (16 CTAs per unit for the 32^3 volume)
            dx = ops.act_zeros(N2, D2, H2, W2, Cc, dev)
            dgam = torch.full((Cc,), 5.0, device=dev); dbet = torch.full((Cc,), 5.0, device=dev)
            ops.gn_backward(xa2, dya2, mean2, rstd2, gamma, beta, dx, dgam, dbet, ws, lrelu=True)
            tag = "C=%d %dx(%d,%d,%d) form %d" % (Cc, N2, D2, H2, W2, ops._lib.lib().b200_gn_backward_form(N2, D2, H2, W2, Cc))
            report("gn_backward dx " + tag, ops.act_to_ncdhw(dx), xr.grad, tol_rel=2e-2)
            report("gn_backward dx vs 3-kernel " + tag, ops.act_to_ncdhw(dx), res["separate"][0], tol_rel=1e-2)
            report("gn_backward dgamma " + tag, dgam, gr.grad, tol_rel=2e-3)
            report("gn_backward dbeta " + tag, dbet, br.grad, tol_rel=2e-3)
            report("gn_backward halo untouched " + tag, ops.act_outside_absmax(dx).view(1), torch.zeros(1, device=dev), tol_abs=0)
            os.environ.pop("B200_GN_BWD_CREG", None)
            report("gn_backward fused == separate dx C=%d" % Cc, res["fused"][0], res["separate"][0], tol_abs=0)
            report("gn_backward fused twice dx C=%d" % Cc, res["fused again"][0], res["fused"][0], tol_abs=0)
            report("gn_backward fused == separate dgamma C=%d" % Cc, res["fused"][1], res["separate"][1], tol_rel=1e-6)
            report("gn_backward fused == separate dbeta C=%d" % Cc, res["fused"][2], res["separate"][2], tol_rel=1e-6)
        # upsample
        xr = xc.clone().requires_grad_(True)
        up = F.leaky_relu(F.interpolate(xr, scale_factor=2, mode="trilinear", align_corners=False), 0.01)
        fine = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cc, dev)
        ops.upsample2x(xa, fine, lrelu=True)
        report("upsample2x+lrelu C=%d" % Cc, ops.act_to_ncdhw(fine), up, tol_rel=1e-2)
        dy = bf(torch.randn_like(up))
        up.backward(dy)
        dco = ops.act_zeros(N, D, H, W, Cc, dev)
        ops.upsample2x_backward(ops.act_from_ncdhw(dy), fine, dco, lrelu=True)
        report("upsample2x backward C=%d" % Cc, ops.act_to_ncdhw(dco), xr.grad, tol_rel=1e-2)
        report("upsample2x halo+guard untouched C=%d" % Cc, ops.act_outside_absmax(fine).view(1), torch.zeros(1, device=dev), tol_abs=0)
    # s2d / d2s / add
    Cc = 16
    xf = bf(torch.randn(N, Cc, 2 * D, 2 * H, 2 * W, device=dev))
    fa = ops.act_from_ncdhw(xf)
    co = ops.act_zeros(N, D, H, W, 8 * Cc, dev)
    ops.space_to_depth(fa, co)
    ref = xf.view(N, Cc, D, 2, H, 2, W, 2).permute(0, 3, 5, 7, 1, 2, 4, 6).reshape(N, 8 * Cc, D, H, W)
    report("space_to_depth", ops.act_to_ncdhw(co), ref, tol_abs=0)
    back = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cc, dev)
    ops.depth_to_space(co, back, residual=fa)
    report("depth_to_space + residual", ops.act_to_ncdhw(back), bf(2 * xf), tol_abs=0)
    s = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cc, dev)
    ops.add(fa, back, s)
    report("add", ops.act_to_ncdhw(s), bf(xf + bf(2 * xf)), tol_abs=0)
    # sigmoid backward + dice
    D2, H2, W2 = 8, 8, 16
    logits = torch.randn(N, 3, D2, H2, W2, device=dev) * 3
    probs = torch.sigmoid(logits)
    target = (torch.rand(N, 3, D2, H2, W2, device=dev) > 0.7).float()
    pr = probs.clone().requires_grad_(True)
    n_, c_ = 2, 3
    inter = (pr * target).view(n_, c_, -1).sum(dim=(0, 2)) + 1e-6
    union = (pr ** 2 + target).view(n_, c_, -1).sum(dim=(0, 2)) + 2e-6
    loss = 1.0 - torch.mean(2.0 * inter / union)
    loss.backward()
    sums = ops.dice_sums(probs, target)
    l = ops.dice_loss(sums, 3, 1.0)
    report("dice loss", l, loss.detach().view(1), tol_abs=1e-5)
    gp = ops.dice_backward(probs, target, sums, torch.ones(1, device=dev), 1.0)
    report("dice backward", gp, pr.grad, tol_rel=1e-3)
    dl = ops.act_zeros(N, D2, H2, W2, 16, dev)
    dbias = torch.empty(3, device=dev)
    ops.sigmoid_backward(gp, probs, dl, dbias)
    refdl = pr.grad * probs * (1 - probs)
    report("sigmoid backward", ops.act_to_ncdhw(dl, 3), bf(refdl), tol_rel=1e-2)
    report("sigmoid backward pad channels zero", ops.act_to_ncdhw(dl)[:, 3:], torch.zeros(N, 13, D2, H2, W2, device=dev), tol_abs=0)
    report("bias grad", dbias, refdl.sum(dim=(0, 2, 3, 4)), tol_rel=2e-3)
    # BCE (loss.py:64-79)
    pr2 = probs.clone().requires_grad_(True)
    bref = -torch.mean(target * torch.log(pr2 + 1e-6) + 1e-2 * (1. - target) * torch.log((1. + 1e-6) - pr2))
    bref.backward()
    bs = ops.bce_sum(probs, target, 1e-2)
    report("bce loss", ops.bce_loss(bs, probs.numel()), bref.detach().view(1), tol_abs=1e-5)
    report("bce backward", ops.bce_backward(probs, target, torch.ones(1, device=dev), 1e-2, probs.numel()), pr2.grad,
           tol_rel=1e-3)


# ----------------------------------------------------------------------------------------------
def _conv3_case(name, N, D, H, W, Cin, Cout, residual=False, lrelu=False, stats=True, seed=0):
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(seed)
    x = bf(torch.randn(N, Cin, D, H, W, device=dev))
    w = bf(torch.randn(Cout, Cin, 3, 3, 3, device=dev) / (Cin * 27) ** 0.5)
    ref = F.conv3d(x, w, padding=1)
    xa = ops.act_from_ncdhw(x)
    desc = ops.conv_desc(ops.MODE_K3, N, D, H, W, ops.pad16(Cin), ops.pad16(Cout))
    packed = ops.conv_pack_weight(desc, ops.W_FWD, w, K_real=Cin, N_real=Cout)
    out = ops.act_zeros(N, D, H, W, ops.pad16(Cout), dev)
    ctas = ops.conv_ctas(desc)
    st = torch.full((ctas, N, 16), 7.0, device=dev) if stats else None
    ra = None
    if residual:
        r = bf(torch.randn(N, Cout, D, H, W, device=dev))
        ra = ops.act_from_ncdhw(r)
        ref = ref + r
    if lrelu:
        ref = F.leaky_relu(ref, 0.01)
    ops.conv_run(desc, xa, packed, out, residual=ra, lrelu=lrelu, stats=st)
    torch.cuda.synchronize()
    ok = report("conv3 %s" % name, ops.act_to_ncdhw(out, Cout), ref, tol_rel=1.5e-2, extra="ctas=%d" % ctas)
    report("conv3 %s halo+guard stay zero" % name, ops.act_outside_absmax(out).view(1), torch.zeros(1, device=dev), tol_abs=0)
    if stats and not residual and not lrelu and Cout % 8 == 0:
        g = F.conv3d(x, w, padding=1).view(N, 8, -1)
        s = st.sum(0)
        report("conv3 %s GN sum" % name, s[:, :8], g.sum(-1), tol_abs=2e-3 * g.abs().sum(-1).max().item())
        report("conv3 %s GN sumsq" % name, s[:, 8:], (g * g).sum(-1), tol_rel=5e-3)
        if Cout % 16 == 0:
            # one-launch form (b200_conv_run_gn: the conv's last CTA finishes the statistics) == conv + gn_finalize,
            # bit for bit; three launches in a row reuse the same ticket word
            mean = torch.empty(N * 8, device=dev); rstd = torch.empty(N * 8, device=dev)
            ops.gn_finalize(st.view(-1), ctas, N, Cout, D, H, W, mean, rstd)
            ticket = torch.zeros(16, dtype=torch.int32, device=dev)
            for rep in range(3):
                out2 = ops.act_zeros(N, D, H, W, Cout, dev)
                st2 = torch.full((ctas, N, 16), -3.0, device=dev)
                m2 = torch.full((N * 8,), 9.0, device=dev); r2 = torch.full((N * 8,), 9.0, device=dev)
                ops.conv_run_gn(desc, xa, packed, out2, st2.view(-1), m2, r2, ticket)
                torch.cuda.synchronize()
            report("conv3 %s fused GN mean (bit-exact)" % name, m2, mean, tol_abs=0)
            report("conv3 %s fused GN rstd (bit-exact)" % name, r2, rstd, tol_abs=0)
            report("conv3 %s fused GN output" % name, out2.t, out.t, tol_abs=0)
            report("conv3 %s fused GN ticket left zero" % name, ticket.float(), torch.zeros(16, device=dev), tol_abs=0)
    return ok


def group_conv3():
    _conv3_case("tiny 16->16 8^3", 1, 8, 8, 8, 16, 16)
    _conv3_case("16->16 2x(6,10,20)", 2, 6, 10, 20, 16, 16)
    _conv3_case("4->16 (conv_input) 16^3", 1, 16, 16, 16, 4, 16)
    _conv3_case("32->32 2x(8,12,16)", 2, 8, 12, 16, 32, 32)
    _conv3_case("64->64 (8,8,16)", 1, 8, 8, 16, 64, 64)
    _conv3_case("128->128 (4,8,8)", 1, 4, 8, 8, 128, 128)
    _conv3_case("128->128 2x(16^3)", 2, 16, 16, 16, 128, 128)
    _conv3_case("16->16 res+lrelu", 1, 8, 8, 24, 16, 16, residual=True, lrelu=True, stats=False)
    _conv3_case("16->16 32^3", 1, 32, 32, 32, 16, 16)
    _conv3_case("16->16 (8,8,128) long lines", 1, 8, 8, 128, 16, 16)
    _conv3_case("32->32 (4,64,64)", 1, 4, 64, 64, 32, 32)
    # wide lines: band-marching kernel (kd and kh folded), full / partial bands and window columns
    _conv3_case("16->16 band (5,16,128)", 1, 5, 16, 128, 16, 16)
    _conv3_case("16->16 band 2x(3,14,110)", 2, 3, 14, 110, 16, 16)
    _conv3_case("16->16 band (3,8,240) two window columns", 1, 3, 8, 240, 16, 16)
    _conv3_case("4->16 band (conv_input) (4,16,128)", 1, 4, 16, 128, 4, 16)
    _conv3_case("16->16 band res+lrelu (4,24,128)", 1, 4, 24, 128, 16, 16, residual=True, lrelu=True, stats=False)
    _conv3_case("16->16 band 2x(20,40,128) many CTAs", 2, 20, 40, 128, 16, 16)
    # sigmoid epilogue (conv_output)
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(3)
    for (N, D, H, W) in ((2, 8, 8, 16), (1, 4, 16, 128)):
        x = bf(torch.randn(N, 16, D, H, W, device=dev))
        w = bf(torch.randn(3, 16, 3, 3, 3, device=dev) * 0.1)
        b = torch.randn(3, device=dev)
        desc = ops.conv_desc(ops.MODE_K3, N, D, H, W, 16, 16, epi=ops.EPI_SIGMOID)
        packed = ops.conv_pack_weight(desc, ops.W_FWD, w, K_real=16, N_real=3)
        probs = torch.empty(N, 3, D, H, W, device=dev); logits = torch.empty_like(probs)
        ops.conv_run(desc, ops.act_from_ncdhw(x), packed, None, bias=b, probs=probs, logits=logits, n_out_real=3)
        ref = F.conv3d(x, w, b, padding=1)
        report("conv_output logits (%d,%d,%d,%d)" % (N, D, H, W), logits, ref, tol_rel=5e-3)
        report("conv_output probs (%d,%d,%d,%d)" % (N, D, H, W), probs, torch.sigmoid(ref), tol_abs=2e-3)


def group_conv1():
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(1)
    N, D, H, W = 2, 6, 8, 12
    for Cin, Cout in ((32, 16), (64, 32), (128, 64), (16, 16)):
        x = bf(torch.randn(N, Cin, D, H, W, device=dev))
        w = bf(torch.randn(Cout, Cin, 1, 1, 1, device=dev) / Cin ** 0.5)
        desc = ops.conv_desc(ops.MODE_K1, N, D, H, W, Cin, Cout)
        packed = ops.conv_pack_weight(desc, ops.W_FWD, w)
        out = ops.act_zeros(N, D, H, W, Cout, dev)
        ops.conv_run(desc, ops.act_from_ncdhw(x), packed, out)
        report("conv1 %d->%d" % (Cin, Cout), ops.act_to_ncdhw(out), F.conv3d(x, w), tol_rel=1.5e-2)
        report("conv1 %d->%d halo+guard stay zero" % (Cin, Cout), ops.act_outside_absmax(out).view(1), torch.zeros(1, device=dev), tol_abs=0)
    # cat conv: two sources
    for Cc in (16, 64):
        a = bf(torch.randn(N, Cc, D, H, W, device=dev)); b = bf(torch.randn(N, Cc, D, H, W, device=dev))
        w = bf(torch.randn(Cc, 2 * Cc, 1, 1, 1, device=dev) / (2 * Cc) ** 0.5)
        desc = ops.conv_desc(ops.MODE_K1, N, D, H, W, Cc, Cc, Cin_b=Cc)
        packed = ops.conv_pack_weight(desc, ops.W_FWD, w)
        out = ops.act_zeros(N, D, H, W, Cc, dev)
        ops.conv_run(desc, ops.act_from_ncdhw(a), packed, out, src_b=ops.act_from_ncdhw(b))
        report("conv1 cat %d+%d->%d" % (Cc, Cc, Cc), ops.act_to_ncdhw(out), F.conv3d(torch.cat([a, b], 1), w), tol_rel=1.5e-2)
        # its data gradients: dskip = W[:, :C]^T dy, dup = W[:, C:]^T dy
        dy = bf(torch.randn(N, Cc, D, H, W, device=dev))
        dd = ops.conv_desc(ops.MODE_K1, N, D, H, W, Cc, Cc)
        for half, nm in ((0, "skip"), (1, "up")):
            pk = ops.conv_pack_weight(dd, ops.W_DGRAD, w, ci_off=half * Cc, K_real=Cc, N_real=Cc)
            o = ops.act_zeros(N, D, H, W, Cc, dev)
            ops.conv_run(dd, ops.act_from_ncdhw(dy), pk, o)
            wt = w[:, half * Cc:(half + 1) * Cc, 0, 0, 0]
            ref = torch.einsum("ncdhw,ck->nkdhw", dy, wt)
            report("conv1 cat dgrad %s C=%d" % (nm, Cc), ops.act_to_ncdhw(o), ref, tol_rel=1.5e-2)
    # k2s2 down conv through space-to-depth, forward and data gradient
    for Cf, Cd in ((16, 32), (32, 64), (64, 128)):
        xf = bf(torch.randn(N, Cf, 2 * D, 2 * H, 2 * W, device=dev))
        w = bf(torch.randn(Cd, Cf, 2, 2, 2, device=dev) / (8 * Cf) ** 0.5)
        s2d = ops.act_zeros(N, D, H, W, 8 * Cf, dev)
        ops.space_to_depth(ops.act_from_ncdhw(xf), s2d)
        desc = ops.conv_desc(ops.MODE_K1, N, D, H, W, 8 * Cf, Cd)
        packed = ops.conv_pack_weight(desc, ops.W_FWD_S2D, w)
        out = ops.act_zeros(N, D, H, W, Cd, dev)
        ops.conv_run(desc, s2d, packed, out)
        report("down k2s2 %d->%d" % (Cf, Cd), ops.act_to_ncdhw(out), F.conv3d(xf, w, stride=2), tol_rel=1.5e-2)
        dy = bf(torch.randn(N, Cd, D, H, W, device=dev))
        dd = ops.conv_desc(ops.MODE_K1, N, D, H, W, Cd, 8 * Cf)
        pk = ops.conv_pack_weight(dd, ops.W_DGRAD_S2D, w)
        o = ops.act_zeros(N, D, H, W, 8 * Cf, dev)
        ops.conv_run(dd, ops.act_from_ncdhw(dy), pk, o)
        fine = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cf, dev)
        ops.depth_to_space(o, fine)
        ref = F.conv_transpose3d(dy, w, stride=2)
        report("down k2s2 dgrad %d<-%d" % (Cf, Cd), ops.act_to_ncdhw(fine), ref, tol_rel=1.5e-2)
        # the same data gradient through the depth-to-space epilogue (EPI_D2S): one launch, written to the fine grid;
        # without a residual it must equal conv + depth_to_space bit for bit, with one it adds the skip gradient in fp32
        d2 = ops.conv_desc(ops.MODE_K1, N, D, H, W, Cd, 8 * Cf, epi=ops.EPI_D2S)
        pk2 = ops.conv_pack_weight(d2, ops.W_DGRAD_S2D, w)
        fine2 = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cf, dev)
        ops.conv_run(d2, ops.act_from_ncdhw(dy), pk2, fine2)
        report("down k2s2 dgrad d2s-epilogue %d<-%d == two-step" % (Cf, Cd), fine2.t, fine.t, tol_abs=0)
        skip = bf(torch.randn(N, Cf, 2 * D, 2 * H, 2 * W, device=dev))
        fine3 = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cf, dev)
        ops.conv_run(d2, ops.act_from_ncdhw(dy), pk2, fine3, residual=ops.act_from_ncdhw(skip))
        report("down k2s2 dgrad d2s-epilogue + skip %d<-%d" % (Cf, Cd), ops.act_to_ncdhw(fine3), ref + skip, tol_rel=1.5e-2)
        report("down k2s2 dgrad d2s-epilogue halo+guard stay zero %d" % Cf, ops.act_outside_absmax(fine3).view(1),
               torch.zeros(1, device=dev), tol_abs=0)
        # the opt-in 32-byte form of the scatter (B200_D2S_V8=1) against the default 16-byte form: the same bits
        os.environ["B200_D2S_V8"] = "1"
        fine4 = ops.act_zeros(N, 2 * D, 2 * H, 2 * W, Cf, dev)
        ops.conv_run(d2, ops.act_from_ncdhw(dy), pk2, fine4, residual=ops.act_from_ncdhw(skip))
        os.environ.pop("B200_D2S_V8", None)
        report("down k2s2 dgrad d2s-epilogue 32-byte == 16-byte form %d" % Cf, fine3.t, fine4.t, tol_abs=0)


def group_dgrad():
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(2)
    for (N, D, H, W, Cin, Cout) in ((1, 8, 8, 8, 16, 16), (2, 6, 10, 12, 32, 32), (1, 4, 8, 8, 64, 64),
                                   (1, 4, 4, 8, 128, 128), (1, 8, 8, 16, 16, 3), (1, 8, 8, 16, 4, 16),
                                   (1, 4, 16, 128, 16, 16), (1, 3, 8, 128, 16, 3)):
        w = bf(torch.randn(Cout, Cin, 3, 3, 3, device=dev) / (Cin * 27) ** 0.5)
        dy = bf(torch.randn(N, Cout, D, H, W, device=dev))
        ref = F.conv_transpose3d(dy, w, padding=1)
        desc = ops.conv_desc(ops.MODE_K3, N, D, H, W, ops.pad16(Cout), ops.pad16(Cin))
        pk = ops.conv_pack_weight(desc, ops.W_DGRAD, w, K_real=Cout, N_real=Cin)
        o = ops.act_zeros(N, D, H, W, ops.pad16(Cin), dev)
        ops.conv_run(desc, ops.act_from_ncdhw(dy), pk, o)
        report("dgrad3 %d<-%d (%d,%d,%d,%d)" % (Cin, Cout, N, D, H, W), ops.act_to_ncdhw(o, Cin), ref, tol_rel=1.5e-2)


def group_wgrad():
    import torch
    import torch.nn.functional as F
    from brats2019_b200 import ops
    dev = "cuda"
    torch.manual_seed(4)
    # 16 <-> 16 channels with W % 16 == 0 take the line-marching kernel (wgrad_line.cuh, the default), or the older
    # marching kernel (wgrad_march.cuh, opt-in with B200_WGRAD_MARCH=1), or the linear-row kernel
    # (B200_NO_WGRAD_LINE=1); those shapes are checked in all three forms: odd D, H not a multiple of the band,
    # batch 2, the real level-0 width 128, 3 and 4 real channels (conv_output / conv_input), one-line volumes
    for (N, D, H, W, Cin, Cout) in ((1, 8, 8, 8, 16, 16), (2, 6, 10, 12, 16, 16), (2, 6, 10, 12, 32, 32),
                                   (1, 6, 8, 8, 64, 64), (2, 4, 4, 8, 128, 128), (1, 8, 8, 16, 16, 3),
                                   (1, 8, 8, 16, 4, 16), (1, 16, 16, 64, 16, 16), (2, 5, 7, 32, 16, 16),
                                   (1, 3, 6, 128, 16, 16), (2, 12, 9, 48, 16, 16), (3, 1, 1, 16, 16, 16),
                                   (1, 20, 37, 128, 16, 16), (2, 7, 16, 144, 16, 16),
                                   # pair form of the line kernel: band heights 8 / 6 / 4 / 2, partial last band, 3 and 4 real channels
                                   (2, 9, 24, 128, 16, 16), (1, 7, 14, 128, 16, 16), (1, 5, 20, 64, 4, 16), (2, 6, 12, 128, 16, 3),
                                   (1, 4, 2, 32, 16, 16), (2, 10, 22, 128, 4, 16),
                                   # 32 <-> 32 channels with W % 16 == 0: the line kernel's three-MMA form (and the linear one)
                                   (1, 8, 8, 16, 32, 32), (2, 5, 7, 32, 32, 32), (2, 9, 12, 64, 32, 32), (1, 3, 6, 48, 32, 32),
                                   (3, 1, 1, 16, 32, 32), (1, 11, 30, 64, 32, 32)):
        x = bf(torch.randn(N, Cin, D, H, W, device=dev))
        dy = bf(torch.randn(N, Cout, D, H, W, device=dev))
        w = torch.zeros(Cout, Cin, 3, 3, 3, device=dev, requires_grad=True)
        F.conv3d(x, w, padding=1).backward(dy)
        desc = ops.wgrad_desc(0, N, D, H, W, ops.pad16(Cout), ops.pad16(Cin))
        # "line" = the default line kernel (one dY line per MMA), "line2" = its opt-in pair mode (two lines per MMA where H is even)
        forms = ("line", "line2", "march", "linear") if (ops.pad16(Cin) == 16 and ops.pad16(Cout) == 16 and W % 16 == 0) else \
            ("line", "linear") if (Cin == 32 and Cout == 32 and W % 16 == 0) else ("linear",)
        for form in forms:
            os.environ["B200_WGRAD_MARCH"] = "1" if form == "march" else "0"
            os.environ["B200_NO_WGRAD_LINE"] = "0" if form in ("line", "line2") else "1"
            os.environ["B200_WGL_PAIR"] = "7" if form == "line2" else "0"
            os.environ["B200_WGRAD_LINE32"] = "1" if form == "line" else "0"      # (the 32-channel line form is opt-in)
            g = torch.full((Cout, Cin, 3, 3, 3), 5.0, device=dev)
            ops.wgrad_run(desc, ops.act_from_ncdhw(dy), ops.act_from_ncdhw(x), g, ops.G_K3)
            report("wgrad3 %dx%d (%d,%d,%d,%d) %s" % (Cout, Cin, N, D, H, W, form), g, w.grad, tol_rel=1e-2)
            if form in ("march", "line", "line2"):      # accumulate into an existing gradient
                g2 = g.clone()
                ops.wgrad_run(desc, ops.act_from_ncdhw(dy), ops.act_from_ncdhw(x), g2, ops.G_K3, accumulate=True)
                report("wgrad3 %dx%d (%d,%d,%d,%d) %s accumulate" % (Cout, Cin, N, D, H, W, form), g2, 2 * w.grad, tol_rel=1e-2)
        os.environ.pop("B200_WGRAD_MARCH", None)
        os.environ.pop("B200_NO_WGRAD_LINE", None)
        os.environ.pop("B200_WGRAD_LINE32", None)
        os.environ.pop("B200_WGL_PAIR", None)
    N, D, H, W = 2, 6, 8, 12
    for Cin, Cout in ((32, 16), (128, 64), (16, 16), (64, 128)):
        x = bf(torch.randn(N, Cin, D, H, W, device=dev)); dy = bf(torch.randn(N, Cout, D, H, W, device=dev))
        ref = torch.einsum("nkdhw,ncdhw->kc", dy, x).view(Cout, Cin, 1, 1, 1)
        desc = ops.wgrad_desc(1, N, D, H, W, Cout, Cin)
        g = torch.zeros(Cout, Cin, 1, 1, 1, device=dev)
        ops.wgrad_run(desc, ops.act_from_ncdhw(dy), ops.act_from_ncdhw(x), g, ops.G_K1)
        report("wgrad1 %dx%d" % (Cout, Cin), g, ref, tol_rel=1e-2)
    # cat conv halves written into one (C, 2C) gradient
    Cc = 32
    a = bf(torch.randn(N, Cc, D, H, W, device=dev)); b = bf(torch.randn(N, Cc, D, H, W, device=dev))
    dy = bf(torch.randn(N, Cc, D, H, W, device=dev))
    ref = torch.einsum("nkdhw,ncdhw->kc", dy, torch.cat([a, b], 1)).view(Cc, 2 * Cc, 1, 1, 1)
    g = torch.zeros(Cc, 2 * Cc, 1, 1, 1, device=dev)
    desc = ops.wgrad_desc(1, N, D, H, W, Cc, Cc)
    ops.wgrad_run(desc, ops.act_from_ncdhw(dy), ops.act_from_ncdhw(a), g, ops.G_K1, ci_off=0)
    ops.wgrad_run(desc, ops.act_from_ncdhw(dy), ops.act_from_ncdhw(b), g, ops.G_K1, ci_off=Cc)
    report("wgrad1 cat halves", g, ref, tol_rel=1e-2)
    # k2s2 through s2d
    for Cf, Cd in ((16, 32), (64, 128)):
        xf = bf(torch.randn(N, Cf, 2 * D, 2 * H, 2 * W, device=dev))
        dy = bf(torch.randn(N, Cd, D, H, W, device=dev))
        w = torch.zeros(Cd, Cf, 2, 2, 2, device=dev, requires_grad=True)
        F.conv3d(xf, w, stride=2).backward(dy)
        s2d = ops.act_zeros(N, D, H, W, 8 * Cf, dev)
        ops.space_to_depth(ops.act_from_ncdhw(xf), s2d)
        desc = ops.wgrad_desc(1, N, D, H, W, Cd, 8 * Cf)
        g = torch.zeros(Cd, Cf, 2, 2, 2, device=dev)
        ops.wgrad_run(desc, ops.act_from_ncdhw(dy), s2d, g, ops.G_S2D)
        report("wgrad k2s2 %dx%d" % (Cd, Cf), g, w.grad, tol_rel=1e-2)


GROUPS = {"ew": group_ew, "conv3": group_conv3, "conv1": group_conv1, "dgrad": group_dgrad, "wgrad": group_wgrad}


def run_group(name):
    import torch
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        GROUPS[name]()
        torch.cuda.synchronize()
    except Exception:
        traceback.print_exc()
        print("FAIL group %s raised" % name, flush=True)
        RESULTS.append(False)
    n_fail = sum(1 for r in RESULTS if not r)
    print("GROUP %s: %d checks, %d failed" % (name, len(RESULTS), n_fail), flush=True)
    return n_fail


def main():
    args = sys.argv[1:]
    if args and args[0] == "--child":
        sys.exit(1 if run_group(args[1]) else 0)
    groups = args or list(GROUPS)
    failed = []
    for g in groups:
        print("==== group %s ====" % g, flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", g], timeout=600)
            if r.returncode != 0:
                failed.append(g)
        except subprocess.TimeoutExpired:
            print("FAIL group %s timed out" % g, flush=True)
            failed.append(g)
    print("OPCHECK SUMMARY: failed groups: %s" % (failed or "none"), flush=True)
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
