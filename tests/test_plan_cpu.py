"""CPU: replay the launch plans of the CUDA kernels in numpy/torch and check the ADDRESSING
(tile decode, halo window, tap row shifts, band/fold stacking, job split) against
torch.nn.functional on small volumes.  No GPU and no compute call into the library: only the
host-side planners (b200_*_plan_debug) run.  What the hardware does with the descriptors is
covered by the `-m gpu` tests; this pins the index arithmetic they are fed.
"""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brats2019_b200 import _lib
from brats2019_b200._lib import ConvDesc, WgradDesc

CONV_FIELDS = ("BD MB TR Q0 QN tiles_q tiles_d num_tiles whole n_jobs KG KGa KC NTG TG x_stages w_stages "
               "x_stage_bytes w_stage_bytes x_plane_bytes SRp nslices halo_rows tmem_cols smem ctas Wp SS fold RB").split()
WGRAD_FIELDS = ("KT XR nband_loaded CoC CiC nfold nacc M Nmma n_jobs splits stages_per_split y_planes "
                "x_planes y_plane_bytes x_plane_bytes stage_bytes stage_tx_bytes stages tmem_cols smem grid banded "
                "folded accs Wp SS nh kw_acc").split()


def conv_plan(mode, N, D, H, W, Cin_a, Cout, Cin_b=0, epi=0):
    d = ConvDesc(mode, epi, N, D, H, W, Cin_a, Cin_b, Cout)
    out = (C.c_int * 96)()
    assert _lib.lib().b200_conv_plan_debug(C.byref(d), out, 96) == 0, _lib.lib().b200_last_error()
    p = {k: out[i] for i, k in enumerate(CONV_FIELDS)}
    p["tap_off"] = [out[len(CONV_FIELDS) + i] for i in range(27)]
    return p


def wgrad_plan(mode, N, D, H, W, Cout, Cin):
    d = WgradDesc(mode, N, D, H, W, Cout, Cin)
    out = (C.c_int * 256)()
    assert _lib.lib().b200_wgrad_plan_debug(C.byref(d), out, 256) == 0, _lib.lib().b200_last_error()
    p = {k: out[i] for i, k in enumerate(WGRAD_FIELDS)}
    nv = len(WGRAD_FIELDS)
    p["jobs"] = [tuple(out[nv + 4 * j + i] for i in range(4)) for j in range(p["n_jobs"])]
    return p


def padded_rows(x):
    """(N,C,D,H,W) -> [rows][C] zero-halo padded, linear row order of common.cuh::Vol::row."""
    N, Cc, D, H, W = x.shape
    a = torch.zeros(N, D + 2, H + 2, W + 2, Cc)
    a[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1)
    return a.reshape(-1, Cc)


def fetch_rows(rows, start, count, guard=None):
    """Bulk-copy semantics with zero guard rows: rows outside [0, total) read as zero, and must lie
    inside the guard the layout reserves (b200_act_guard_rows)."""
    total = rows.shape[0]
    if guard is not None:
        assert start >= -guard and start + count <= total + guard, (start, count, total, guard)
    out = torch.zeros(count, rows.shape[1])
    lo, hi = max(start, 0), min(start + count, total)
    if hi > lo:
        out[lo - start:hi - start] = rows[lo:hi]
    return out


@pytest.mark.parametrize("shape", [(1, 4, 4, 4), (2, 5, 6, 9), (1, 8, 8, 8), (1, 3, 16, 40), (2, 16, 16, 16)])
@pytest.mark.parametrize("cout", [16, 32, 128])
def test_conv3_addressing(shape, cout):
    N, D, H, W = shape
    Cin = 16
    p = conv_plan(0, N, D, H, W, Cin, cout)
    torch.manual_seed(0)
    x = torch.randn(N, Cin, D, H, W)
    w = torch.randn(4, Cin, 3, 3, 3)              # 4 real output channels are enough to pin addressing
    ref = F.conv3d(x, w, padding=1)
    rows = padded_rows(x)
    Wp, SS, Dp = p["Wp"], p["SS"], D + 2
    guard = _lib.lib().b200_act_guard_rows(D, H, W)
    assert _lib.lib().b200_act_plane_rows(N, D, H, W) == rows.shape[0] + 2 * guard
    out = torch.zeros(N, 4, D, H, W)
    written = torch.zeros(N, D, H, W, dtype=torch.int32)
    wt = w.reshape(4, Cin, 27)
    assert p["num_tiles"] == (N * p["tiles_d"]) * p["tiles_q"]
    for t in range(p["num_tiles"]):
        tq, r = t % p["tiles_q"], t // p["tiles_q"]
        td, n = r % p["tiles_d"], r // p["tiles_d"]
        d0, q0 = td * p["BD"], p["Q0"] + tq * p["TR"]
        # stage image: nslices x SRp rows
        plane = torch.zeros(p["nslices"] * p["SRp"], Cin)
        for s in range(p["nslices"]):
            dpi = (0 if p["whole"] else d0 + 1) - 1 + s
            row0 = (n * Dp + dpi) * SS + q0 - p["halo_rows"]
            plane[s * p["SRp"]:(s + 1) * p["SRp"]] = fetch_rows(rows, row0, p["SRp"], guard)
        fold, RB = p["fold"], p["RB"]
        for run in range(p["BD"] * p["MB"]):
            dz, mb = run // p["MB"], run % p["MB"]
            if fold:
                # N = (kw, co): 9 (kd,kh) MMAs, then out[m] = P[m-1][kw0] + P[m][kw1] + P[m+1][kw2]
                P = torch.zeros(128, 3, 4)
                for t9 in range(9):
                    a0 = dz * p["SRp"] + mb * RB + p["tap_off"][t9]
                    assert a0 + 128 <= plane.shape[0], "A operand would read past the stage plane"
                    for kw in range(3):
                        P[:, kw] += plane[a0:a0 + 128] @ wt[:, :, t9 * 3 + kw].T
                acc = torch.zeros(128, 4)
                acc[1:127] = P[0:126, 0] + P[1:127, 1] + P[2:128, 2]
            else:
                acc = torch.zeros(128, 4)
                for tap in range(27):
                    a0 = dz * p["SRp"] + mb * 128 + p["tap_off"][tap]
                    assert a0 + 128 <= plane.shape[0], "A operand would read past the stage plane"
                    acc += plane[a0:a0 + 128] @ wt[:, :, tap].T
            for m in range(128):
                if fold and not (1 <= m <= RB):
                    continue
                q = q0 + mb * RB + m - (1 if fold else 0)
                dpo = (0 if p["whole"] else d0 + 1) + dz
                dq, r2 = divmod(q, SS)
                hp, wp = divmod(r2, Wp)
                dp = dpo + dq
                valid = q < p["Q0"] + p["QN"] and 1 <= dp <= D and 1 <= hp <= H and 1 <= wp <= W
                if valid:
                    orow = (n * Dp + dpo) * SS + q
                    assert orow == ((n * Dp + dp) * (H + 2) + hp) * Wp + wp
                    out[n, :, dp - 1, hp - 1, wp - 1] = acc[m]
                    written[n, dp - 1, hp - 1, wp - 1] += 1
    assert (written == 1).all(), "every interior voxel must be produced exactly once"
    np.testing.assert_allclose(out.numpy(), ref.numpy(), atol=1e-4)
    assert p["smem"] <= 227 * 1024 and p["tmem_cols"] <= 512
    assert 2 * p["BD"] * p["MB"] * (3 * cout if p["fold"] else min(cout, 256)) <= p["tmem_cols"]
    assert p["fold"] == (1 if cout in (16, 32) else 0)


def test_conv1_plan_covers_all_rows():
    p = conv_plan(1, 2, 6, 8, 12, 32, 16, Cin_b=32)
    total = 2 * 8 * 10 * 14
    assert p["num_tiles"] * p["TR"] >= total and (p["num_tiles"] - 1) * p["TR"] < total
    assert p["KG"] == 2 * p["KGa"] and p["KC"] * p["KGa"] == 32 and p["SRp"] == p["TR"]
    p = conv_plan(1, 1, 4, 4, 4, 64, 512)
    assert p["n_jobs"] == 2 and p["tmem_cols"] == 512


@pytest.mark.parametrize("cfg", [(16, 16), (32, 32), (64, 64), (128, 128)])
def test_wgrad3_addressing(cfg):
    Cout, Cin = cfg
    N, D, H, W = 2, 3, 4, 5
    p = wgrad_plan(0, N, D, H, W, Cout, Cin)
    torch.manual_seed(1)
    co_r, ci_r = 3, 2          # real channels used for the check (others zero) keeps it fast
    x = torch.zeros(N, Cin, D, H, W); x[:, :ci_r] = torch.randn(N, ci_r, D, H, W)
    dy = torch.zeros(N, Cout, D, H, W); dy[:, :co_r] = torch.randn(N, co_r, D, H, W)
    w = torch.zeros(Cout, Cin, 3, 3, 3, requires_grad=True)
    F.conv3d(x, w, padding=1).backward(dy)
    ref = w.grad[:co_r, :ci_r]
    X, Y = padded_rows(x), padded_rows(dy)
    total = X.shape[0]
    Wp, SS, KT = p["Wp"], p["SS"], p["KT"]
    guard = _lib.lib().b200_act_guard_rows(D, H, W)
    M, Nm, nacc, nh = p["M"], p["Nmma"], p["nacc"], p["nh"]
    # dual plans (nh == 2) stack two row ranges of a stage in M and N; the replay keeps the logical tile
    # [bands*Cout][folds*Cin] and adds the two ranges (the physical interleave is checked on the GPU)
    Ml, Nl = M // nh, Nm // nh
    partial = torch.zeros(p["n_jobs"], p["splits"], nacc, Ml, Nl)
    stage_rows = KT * nh
    for job, (jkd, jkh, jkw, jx) in enumerate(p["jobs"]):
        for split in range(p["splits"]):
            first = split * p["stages_per_split"] * stage_rows
            nst = min(p["stages_per_split"], max(0, -(-(total - first) // stage_rows)))
            for i in range(nst):
                for h in range(nh):
                    r0 = first + i * stage_rows + h * KT
                    A = torch.zeros(KT, Ml)        # [k][m]; unloaded bands stay zero here (garbage on HW)
                    for b in range(p["nband_loaded"]):
                        kd = (b - 1) if p["nband_loaded"] > 1 else jkd
                        A[:, b * Cout:(b + 1) * Cout] = fetch_rows(Y, r0 - kd * SS, KT, guard)
                    XR = p["XR"]
                    Bm = torch.zeros(XR, Nl)
                    # one of kh/kw is the fold axis (copies of X), the other the accumulator axis (start shifts)
                    fsh, ash = (Wp, 1) if p["kw_acc"] else (1, Wp)
                    jf, ja = (jkh, jkw) if p["kw_acc"] else (jkw, jkh)
                    for f in range(p["nfold"]):
                        ft = (f - 1) if p["nfold"] > 1 else jf
                        xr = r0 + ft * fsh + (-1 if nacc > 1 else ja) * ash
                        Bm[:, f * Cin:(f + 1) * Cin] = fetch_rows(X, xr, XR, guard)[:, jx:jx + Cin]
                    for t in range(nacc):
                        xrow = t * ash if nacc > 1 else 0
                        assert xrow + KT <= XR
                        partial[job, split, t] += A.T @ Bm[xrow:xrow + KT]
    red = partial.sum(1)
    got = torch.zeros(co_r, ci_r, 3, 3, 3)
    for co in range(co_r):
        for ci in range(ci_r):
            for tp in range(27):
                kd, kh, kw = tp // 9, (tp // 3) % 3, tp % 3
                tf, ta = (kh, kw) if p["kw_acc"] else (kw, kh)
                job = ((0 if p["banded"] else kd) * (1 if p["accs"] else 3) + (0 if p["accs"] else ta)) * \
                      (1 if p["folded"] else 3) + (0 if p["folded"] else tf)
                row = (kd * Cout if p["banded"] else 0) + co
                col = (tf * Cin if p["folded"] else 0) + ci
                t = ta if p["accs"] else 0
                got[co, ci, kd, kh, kw] = red[job, t, row, col]
    np.testing.assert_allclose(got.numpy(), ref.numpy(), atol=1e-3)
    assert p["smem"] <= 227 * 1024 and p["tmem_cols"] <= 512 and p["nacc"] * p["Nmma"] <= p["tmem_cols"]
    assert p["grid"] == p["splits"] * p["n_jobs"]


def test_library_exports_every_declared_symbol():
    import os
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "brats_b200.h")).read()
    declared = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    L = _lib.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert declared == set(_lib.EXPORTED_SYMBOLS)


def test_fastdiv_formula_is_exact_for_planner_divisors():
    """Kernels divide by W, W*C/8, Wp and SS = Hp*Wp with umulhi(n, ceil(2^(31+s)/d)) >> (s-1);
    operands stay below 2^24 (rows of one padded sample, vectors of one CTA)."""
    divs = {2, 3, 5, 6, 7, 10, 12, 16, 20, 24, 30, 32, 40, 48, 60, 64, 80, 96, 120, 128, 160, 192, 240, 256, 320, 480,
            512, 640, 960, 1280, 1920, 2560, 3840}
    for side in (16, 24, 32, 48, 64, 96, 128, 144, 160, 192, 240, 20, 40, 80, 30, 60, 120, 18, 36, 72):
        divs.add(side + 2)
        for other in (16, 32, 64, 128, 144, 160, 192, 240, 20, 40, 80, 30, 60, 120):
            divs.add((side + 2) * (other + 2))
    rng = np.random.default_rng(0)
    for d in sorted(divs):
        sh = 0
        while (1 << sh) < d:
            sh += 1
        m = ((1 << (31 + sh)) + d - 1) // d
        assert m < 2 ** 32
        # the quotient can only be wrong next to a multiple of d: test every k*d-1, k*d, k*d+1 below 2^24
        k = np.arange(0, (1 << 24) // d + 1, dtype=np.uint64) * np.uint64(d)
        n = np.concatenate([k, k + np.uint64(1), np.maximum(k, np.uint64(1)) - np.uint64(1),
                            rng.integers(0, 1 << 24, 4096).astype(np.uint64)])
        n = n[n < (1 << 24)]
        q = ((n * np.uint64(m)) >> np.uint64(32)) >> np.uint64(sh - 1)
        assert (q == n // np.uint64(d)).all(), d


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No fallback of any kind: without the built CUDA library the first native call raises."""
    import pytest as _pytest
    from brats2019_b200 import _lib as L
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", str(tmp_path / "libbrats_b200.so"))
    with _pytest.raises(RuntimeError, match="no fallback"):
        L.lib()


def test_cpu_tensor_is_rejected_not_computed():
    """The drop-in module refuses CPU inputs instead of falling back to ATen (model.py::UNet.forward)."""
    import contextlib
    import sys
    import pytest as _pytest
    import torch
    import brats2019_b200 as B
    with contextlib.redirect_stdout(sys.stderr):
        m = B.UNet(**B.DEFAULT_CFG)
    with _pytest.raises(RuntimeError, match="no CPU"):
        m([torch.zeros(1, 4, 16, 16, 16)])
    with _pytest.raises(RuntimeError, match="CUDA only"):
        B.Dice_loss_joint()([torch.zeros(1, 3, 8, 8, 8)], [torch.zeros(1, 3, 8, 8, 8)])


# ---------------------------------------------------------------------------------------------------------------
# Depth-to-space epilogue of the k2 s2 data gradient (conv_gemm.cuh, EPI_D2S): the (row block, column group) items a
# tile's accumulators are cut into are dealt round-robin to the three epilogue groups.  Replay of the item -> (fine
# chunk, fine row) arithmetic against the definition of space-to-depth (elementwise.cuh: s2d channel
# ((kd*2+kh)*2+kw)*Cf + c of coarse voxel (d,h,w) = fine[2d+kd][2h+kh][2w+kw][c]) - every fine (voxel, chunk) written
# exactly once, by exactly one group, and only interior voxels.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("Cf", [16, 32, 64])
@pytest.mark.parametrize("dims", [(1, 2, 3, 4), (2, 3, 2, 6)])
def test_d2s_epilogue_item_dealing_covers_the_fine_grid_once(Cf, dims):
    N, D, H, W = dims                                   # coarse volume
    groups = 3                                          # kEpiGroups
    Cout = 8 * Cf                                       # GEMM N: (tap8, fine channel)
    CO = 128 if Cout == 128 else 256                    # columns per CTA job (conv_nmma)
    n_jobs = Cout // CO
    R = 512 // (2 * CO)                                 # row blocks per tile: two accumulator stages in 512 TMEM columns
    NCG = CO // 16
    sh = (Cf // 8).bit_length() - 1                     # d2s_sh = log2(Cf / 8)
    Wp, Hp, Dp = W + 2, H + 2, D + 2
    SS = Hp * Wp
    total_rows = N * Dp * SS
    fWp, fHp, fDp = 2 * W + 2, 2 * H + 2, 2 * D + 2
    fslice = fHp * fWp
    TR = 128 * R
    written = {}                                        # (fine chunk, fine padded row) -> (job, group)
    for job in range(n_jobs):
        for t in range((total_rows + TR - 1) // TR):
            for grp in range(groups):
                for item in range(grp, R * NCG, groups):
                    r, c0 = item // NCG, (item % NCG) * 16
                    for m in range(128):
                        orow = t * TR + r * 128 + m
                        if orow >= total_rows:
                            continue
                        dpf, r2 = divmod(orow, SS)
                        hp, wp = divmod(r2, Wp)
                        nn, dp = divmod(dpf, Dp)
                        if not (1 <= dp <= D and 1 <= hp <= H and 1 <= wp <= W):
                            continue                    # halo rows hold nothing
                        frow0 = ((nn * fDp + (2 * dp - 1)) * fHp + (2 * hp - 1)) * fWp + (2 * wp - 1)
                        chg = (job * CO + c0) >> 3
                        tap8, ch = chg >> sh, chg & ((1 << sh) - 1)
                        row = frow0 + (tap8 >> 2) * fslice + ((tap8 >> 1) & 1) * fWp + (tap8 & 1)
                        for cc in (ch, ch + 1):         # a column group is two 8-channel chunks
                            assert (cc, row) not in written, "fine element written twice"
                            written[(cc, row)] = (job, grp)
                        # definition: GEMM column n = tap8 * Cf + c  <->  fine voxel (2d+kd, 2h+kh, 2w+kw), channel c
                        n0 = job * CO + c0
                        kd, kh, kw = (n0 // Cf) >> 2, ((n0 // Cf) >> 1) & 1, (n0 // Cf) & 1
                        d, h, w = dp - 1, hp - 1, wp - 1
                        want = ((nn * fDp + (2 * d + kd + 1)) * fHp + (2 * h + kh + 1)) * fWp + (2 * w + kw + 1)
                        assert row == want and ch == (n0 % Cf) // 8
    # every interior fine voxel x every fine chunk, nothing else
    assert len(written) == N * (2 * D) * (2 * H) * (2 * W) * (Cf // 8)
    for (cc, row) in written:
        nd, r2 = divmod(row, fslice)
        fh, fw = divmod(r2, fWp)
        fd = nd % fDp
        assert 1 <= fd <= 2 * D and 1 <= fh <= 2 * H and 1 <= fw <= 2 * W and 0 <= cc < Cf // 8
    assert len({g for (_, g) in written.values()}) == min(groups, R * NCG)      # all three groups take part
