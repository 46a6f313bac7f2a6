"""CPU replay of the marching 3x3x3 kernel's control logic (csrc/conv_march.cuh) from the integers
the library's planner exports (b200_march_plan_debug): unit ranges per CTA, segments, the slice ring,
the rotation of the kd taps over the three TMEM accumulator slots, slot zeroing and the validity of
every stored row.  The replay must reproduce F.conv3d exactly (float64) and write every interior
voxel exactly once - for the planner's grid and for other grid sizes (the kernel is grid-agnostic).
No GPU needed: only the planner runs (host code of libbrats_b200.so)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brats2019_b200 import _lib, ops

NAMES = ("MB TR Q0 QN n_strips units KS SRp nslots plane_bytes slot_bytes w_bytes wtile_bytes tmem_cols smem ctas "
         "Wp SS").split()


def march_plan(N, D, H, W, Cin, Cout):
    d = ops.conv_desc(ops.MODE_K3, N, D, H, W, Cin, Cout)
    out = (C.c_int * 32)()
    _lib.check(_lib.lib().b200_march_plan_debug(C.byref(d), out, 32), "b200_march_plan_debug")
    return dict(zip(NAMES, list(out)))


def replay(P, x, w, ctas):
    N, Cin, D, H, W = x.shape
    CO = w.shape[0]
    Wp, SS, MB, TR, Q0, QN, D_ = P["Wp"], P["SS"], P["MB"], P["TR"], P["Q0"], P["QN"], D
    Dp = D + 2
    rows = N * Dp * SS
    guard = SS + 2 * Wp + 1024
    X = np.zeros((rows + guard, Cin))
    xp = np.zeros((N, Dp, H + 2, Wp, Cin))
    xp[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1).numpy()
    X[:rows] = xp.reshape(rows, Cin)
    wn = w.numpy()
    out = np.zeros((rows, CO))
    written = np.zeros(rows, dtype=np.int32)
    units = P["units"]
    steps_total = 0
    for cta in range(ctas):
        u, u_end = units * cta // ctas, units * (cta + 1) // ctas
        acc = np.zeros((MB, 3, 128, CO))           # TMEM: three accumulator slots per block, zero at start
        while u < u_end:
            d0 = u % D_
            t = u // D_
            strip, n = t % P["n_strips"], t // P["n_strips"]
            ln = min(D_ - d0, u_end - u)
            d1 = d0 + ln - 1
            for dpi in range(d0, d1 + 3):
                steps_total += 1
                row0 = (n * Dp + dpi) * SS + strip * TR
                slot_rows = X[row0:row0 + P["SRp"]]
                assert slot_rows.shape[0] == P["SRp"], "slot read runs past the guard rows"
                rot = (4 - dpi % 3) % 3
                bands = [2, 1, 0, 2, 1][rot:rot + 3]
                dpo, s = dpi - 1, (dpi + 2) % 3
                for b in range(MB):
                    for kh in range(3):
                        for kw in range(3):
                            A = slot_rows[128 * b + kh * Wp + kw:128 * b + kh * Wp + kw + 128]
                            for j, kd in enumerate(bands):
                                acc[b, j] += A @ wn[:, :, kd, kh, kw].T
                    vals = acc[b, s].copy()
                    if dpi == d1 + 2:
                        acc[b] = 0
                    else:
                        acc[b, s] = 0
                    if dpi < d0 + 2:
                        continue
                    for m in range(128):
                        q = Q0 + strip * TR + 128 * b + m
                        hp, wp = divmod(q, Wp)
                        if q < Q0 + QN and 1 <= hp <= H and 1 <= wp <= W:
                            r = (n * Dp + dpo) * SS + q
                            out[r] = vals[m]
                            written[r] += 1
            u += ln
    o = out.reshape(N, Dp, H + 2, Wp, CO)
    wr = written.reshape(N, Dp, H + 2, Wp)
    return o, wr, steps_total


@pytest.mark.parametrize("shape,cin,cout", [((2, 6, 20, 24), 16, 16), ((1, 4, 8, 8), 16, 16), ((2, 3, 12, 40), 32, 32),
                                            ((1, 5, 30, 10), 16, 32)])
def test_march_replay_matches_conv3d(shape, cin, cout):
    N, D, H, W = shape
    P = march_plan(N, D, H, W, cin, cout)
    assert P["MB"] * 3 * cout <= P["tmem_cols"] <= 512
    assert P["smem"] <= 227 * 1024 and P["nslots"] >= 3
    assert P["slot_bytes"] == (cin // 8) * P["SRp"] * 16 and P["SRp"] >= P["TR"] + 2 * P["Wp"] + 2
    assert P["units"] == N * P["n_strips"] * D and P["n_strips"] * P["TR"] >= P["QN"]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, cin, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, padding=1).permute(0, 2, 3, 4, 1).numpy()
    for ctas in sorted({P["ctas"], 1, 3, 7}):
        if ctas > P["units"]:
            continue
        o, wr, _ = replay(P, x, w, ctas)
        assert (wr[:, 1:-1, 1:-1, 1:-1] == 1).all(), "interior voxel not written exactly once (ctas=%d)" % ctas
        wr[:, 1:-1, 1:-1, 1:-1] = 0
        assert (wr == 0).all(), "a halo row was written (ctas=%d)" % ctas
        np.testing.assert_allclose(o[:, 1:-1, 1:-1, 1:-1], ref, rtol=1e-9, atol=1e-9)


def test_march_plans_for_bench_shapes():
    """The shapes of the benchmark and of BraTS inference get a marching plan that fits the SM."""
    for (N, S, Cc) in [(2, 128, 16), (2, 64, 32), (1, 128, 16)]:
        P = march_plan(N, S, S, S, Cc, Cc)
        assert P["smem"] <= 227 * 1024 and P["tmem_cols"] <= 512 and P["ctas"] == 148
    P = march_plan(1, 160, 240, 240, 16, 16)
    assert P["smem"] <= 227 * 1024
    P = march_plan(1, 80, 120, 120, 32, 32)
    assert P["smem"] <= 227 * 1024


# ------------------------------------------------------------------------------------------------
# band-marching kernel (csrc/conv_band.cuh): kd and kh folded, accumulators [line slot][slice slot]
# ------------------------------------------------------------------------------------------------
BAND_NAMES = "BH n_bands n_cols units nslots plane_bytes slot_bytes w_bytes wimg_bytes smem ctas Wp SS".split()


def band_plan(N, D, H, W):
    d = ops.conv_desc(ops.MODE_K3, N, D, H, W, 16, 16)
    out = (C.c_int * 32)()
    _lib.check(_lib.lib().b200_band_plan_debug(C.byref(d), out, 32), "b200_band_plan_debug")
    return dict(zip(BAND_NAMES, list(out)))


def replay_band(P, x, w, ctas):
    N, Cin, D, H, W = x.shape
    CO = w.shape[0]
    Wp, SS, BH = P["Wp"], P["SS"], P["BH"]
    Dp, Hp = D + 2, H + 2
    rows = N * Dp * SS
    guard = SS + (BH + 2) * Wp + 1024
    X = np.zeros((rows + guard, Cin))
    xp = np.zeros((N, Dp, Hp, Wp, Cin))
    xp[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1).numpy()
    X[:rows] = xp.reshape(rows, Cin)
    wn = w.numpy()
    out = np.zeros((rows, CO))
    written = np.zeros(rows, dtype=np.int32)
    units, nb, nc = P["units"], P["n_bands"], P["n_cols"]
    for cta in range(ctas):
        u, u_end = units * cta // ctas, units * (cta + 1) // ctas
        acc = np.zeros((BH + 2, 3, 128, CO))
        while u < u_end:
            d0 = u % D
            t = u // D
            col = t % nc
            t //= nc
            band, n = t % nb, t // nb
            ln = min(D - d0, u_end - u)
            d1 = d0 + ln - 1
            for dpi in range(d0, d1 + 3):
                rot = (4 - dpi % 3) % 3
                kds = [2, 1, 0, 2, 1][rot:rot + 3]
                for ji in range(BH + 2):
                    row0 = ((n * Dp + dpi) * Hp + band * BH + ji) * Wp + col * 128
                    seg = X[row0:row0 + 130]
                    assert seg.shape[0] == 130, "line segment read runs past the guard rows"
                    if ji == 0:
                        khs = [(0, 1)]
                    elif ji == BH + 1:
                        khs = [(2, BH)]
                    else:
                        khs = [(2, ji - 1), (1, ji), (0, ji + 1)]
                    for kw in range(3):
                        A = seg[kw:kw + 128]
                        for kh, jo in khs:
                            for s, kd in enumerate(kds):
                                acc[jo, s] += A @ wn[:, :, kd, kh, kw].T
                dpo, s = dpi - 1, (dpi + 2) % 3
                for jo in range(1, BH + 1):
                    vals = acc[jo, s].copy()
                    if dpi == d1 + 2:
                        acc[jo] = 0
                    else:
                        acc[jo, s] = 0
                    hp = band * BH + jo
                    if dpi < d0 + 2 or hp > H:
                        continue
                    for m in range(128):
                        wp = 1 + col * 128 + m
                        if wp <= W:
                            r = ((n * Dp + dpo) * Hp + hp) * Wp + wp
                            out[r] = vals[m]
                            written[r] += 1
            u += ln
    return out.reshape(N, Dp, Hp, Wp, CO), written.reshape(N, Dp, Hp, Wp)


@pytest.mark.parametrize("shape", [(1, 4, 16, 128), (2, 3, 14, 110), (1, 3, 8, 240)])
def test_band_replay_matches_conv3d(shape):
    N, D, H, W = shape
    P = band_plan(N, D, H, W)
    assert P["smem"] <= 227 * 1024 and P["nslots"] >= 3 and (P["BH"] + 2) * 48 <= 512
    assert P["units"] == N * P["n_bands"] * P["n_cols"] * D
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, 16, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(16, 16, 3, 3, 3, generator=g, dtype=torch.float64)
    ref = F.conv3d(x, w, padding=1).permute(0, 2, 3, 4, 1).numpy()
    for ctas in sorted({min(P["ctas"], 5), 1, 3}):
        if ctas > P["units"]:
            continue
        o, wr = replay_band(P, x, w, ctas)
        assert (wr[:, 1:-1, 1:-1, 1:-1] == 1).all(), "interior voxel not written exactly once (ctas=%d)" % ctas
        wr[:, 1:-1, 1:-1, 1:-1] = 0
        assert (wr == 0).all(), "a halo row was written (ctas=%d)" % ctas
        np.testing.assert_allclose(o[:, 1:-1, 1:-1, 1:-1], ref, rtol=1e-9, atol=1e-9)


def test_band_plan_applies_to_level0_shapes_only():
    assert band_plan(2, 128, 128, 128)["ctas"] == 148
    assert band_plan(1, 160, 240, 240)["n_cols"] == 2
    d = ops.conv_desc(ops.MODE_K3, 1, 16, 16, 16, 16, 16)       # narrow lines: plain marching kernel
    out = (C.c_int * 32)()
    assert _lib.lib().b200_band_plan_debug(C.byref(d), out, 32) != 0
