"""CPU replay of the line-marching weight-gradient kernel (brats2019_b200/csrc/wgrad_line.cuh): the planner's
numbers (b200_wgrad_line_plan_debug) drive a numpy emulation of what a CTA does, including its asynchronous
structure - three actors (X producer, dY producer, MMA issuer) that only advance when the kernel's own wait rules allow
it, scheduled in random order with the producers as greedy as the rules let them be:

  * shared-memory ring: slot (3*line + slice) mod R with the first 8 slots mirrored behind the ring; every 9-slot
    window the MMA reads must hold exactly the lines (kh, kd) of the step, i.e. no load may land in a slot that a
    later step still reads (the reuse rule `step (s-3)*nl + min(line+1, nl-1) done`) and no window may wrap;
  * the ring of Ny expanded dY lines: line t lands (three bulk copies per chunk) only after the MMAs of line t - Ny;
  * mbarrier rings: a waiter is never two phases behind the barrier it polls (x_full[64], step_done[16], y_exp[Ny]);
  * operands: the (kw, chunk) copies of a dY line and the K range over the interior voxels of a line;
  * the reduce kernel's index map back to the PyTorch (Cout, Cin, 3, 3, 3) gradient, against torch autograd.
No GPU: what the hardware does with the descriptors is covered by tests/gpu_opcheck.py (wgrad group) and
tests/test_layer_parity_gpu.py."""
import ctypes as C
import random

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brats2019_b200 import _lib
from brats2019_b200._lib import WgradDesc

FIELDS = "LH n_bands units ksteps R Ny Wp Lp smem_x_off smem_y_off smem_bar_off smem grid mirror NB ND NR smem_raw_off nchy nchx pair".split()


def line_plan(N, D, H, W, Cc=16, real_out=0, real_in=0, pair=False):
    """real_out / real_in: PyTorch channel counts of the gradient (<= 8 on one side: that side's upper chunk is skipped).
    pair: True = B200_WGL_PAIR=7 (opt-in: two dY lines per MMA wherever it applies), False = the default (one line per MMA)."""
    import os
    d = WgradDesc(0, N, D, H, W, Cc, Cc)
    out = (C.c_int * 32)()
    old = os.environ.get("B200_WGL_PAIR")
    os.environ["B200_WGL_PAIR"] = "7" if pair else "0"
    try:
        assert _lib.lib().b200_wgrad_line_plan_debug2(C.byref(d), real_out, real_in, out, 32) == 0, _lib.lib().b200_last_error()
    finally:
        if old is None:
            del os.environ["B200_WGL_PAIR"]
        else:
            os.environ["B200_WGL_PAIR"] = old
    return {k: out[i] for i, k in enumerate(FIELDS)}


def padded(x):
    """(N,C,D,H,W) -> (N, D+2, H+2, W+2, C) with the zero halo the activation layout keeps in HBM."""
    N, Cc, D, H, W = x.shape
    a = np.zeros((N, D + 2, H + 2, W + 2, Cc), np.float64)
    a[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1).numpy()
    return a


def segments(p, cta, N, D, H):
    units, grid, nb, LH = p["units"], p["grid"], p["n_bands"], p["LH"]
    u, u_end = units * cta // grid, units * (cta + 1) // grid
    out = []
    while u < u_end:
        d0, band, n = u % D, (u // D) % nb, u // D // nb
        length = min(D - d0, u_end - u)
        out.append(dict(n=n, band=band, d0=d0, len=length, nl=min(LH, H - band * LH)))
        u += length
    return out


def replay_cta(p, segs, Y, X, W, rng):
    """One CTA.  Returns its (3C padded to 64 / 128) x 9C accumulator (C = 16: one N = 144 MMA per K step; C = 32: three
    N = 96 MMAs, one per kh, into the column blocks kh * 96)."""
    R, Ny, NR, NB, ND, MIR, LH = p["R"], p["Ny"], p["NR"], p["NB"], p["ND"], p["mirror"], p["LH"]
    Cy, Cx = 8 * p["nchy"], 8 * p["nchx"]                 # channels of dY / X the kernel loads
    PAIR = p["pair"]
    LN, WIN = (2, 12) if PAIR else (1, 9)        # dY lines per MMA, ring slots per B operand
    assert MIR == WIN - 1
    Mm = 64 if LN * 3 * Cy <= 64 else 128
    ring = [None] * (R + MIR)                    # slot -> (seg index, slice s, line lam) currently stored
    yslot = [None] * Ny                          # slot -> step index whose expanded line it holds
    acc = np.zeros((Mm, WIN * Cx))
    # ---- static schedules of the three actors (exactly the kernel's loops) ----
    xloads, steps = [], []                       # (seg, s, lam, wait_step or None) ; (seg, sd, l, k_need)
    t_base, seg_k0 = 0, 0
    for si, sg in enumerate(segs):
        nl = sg["nl"]
        for s in range(sg["len"] + 2):
            for lam in range(nl + 2):
                wait = None
                if s >= 3:
                    wait = t_base + (s - 3) * nl + min(lam + 1, nl - 1)
                if s == 0 and lam == 0 and t_base > 0:
                    wait = t_base - 1
                xloads.append((si, s, lam, wait))
        for sd in range(1, sg["len"] + 1):
            for l in range(nl):
                steps.append((si, sd, l, seg_k0 + (sd + 1) * (nl + 2) + l + 2))
        t_base += sg["len"] * nl
        seg_k0 += (sg["len"] + 2) * (nl + 2)
    nsteps = len(steps)
    kx = ky = t = 0                              # progress of X producer, dY producer, MMA
    max_ahead_x = 0
    guard = 0
    while t < nsteps:
        guard += 1
        assert guard < 50 * (len(xloads) + 2 * nsteps) + 1000, "deadlock: no actor can advance"
        order = [0, 0, 0, 1, 1, 1, 2]            # producers greedy, MMA lazy: the hardest interleaving for slot reuse
        rng.shuffle(order)
        for who in order:
            if who == 0 and kx < len(xloads):
                si, s, lam, wait = xloads[kx]
                if wait is not None and wait >= t:
                    continue                                     # step `wait` not done yet
                if wait is not None:
                    assert t - 1 - wait < ND, "step_done phase ambiguity for the X producer"
                waited = steps[t - 1][3] + 1 if t > 0 else 0      # loads the MMA warp has certainly waited for
                assert kx - waited < NB, "x_full ring lapped"
                q = (3 * lam + s) % R
                ring[q] = (si, s, lam)
                if q < MIR:
                    ring[q + R] = (si, s, lam)
                kx += 1
            elif who == 1 and ky < nsteps:
                if ky >= Ny and ky - Ny >= t:
                    continue                                     # the MMAs of line ky - Ny still read the slot (step_done)
                if ky >= Ny:
                    assert t - 1 - (ky - Ny) < ND, "step_done phase ambiguity for the dY producer"
                yslot[ky % Ny] = ky
                ky += 1
            elif who == 2 and t < nsteps and PAIR:
                # pair mode: an issue block is up to four lines (two pairs) of one slice; the MMA warp polls `ready >= t + n4`
                si, sd, l, _ = steps[t]
                nl = segs[si]["nl"]
                assert nl % 2 == 0 and l % 4 == 0 and Ny % 2 == 0
                n4 = min(4, nl - l)
                k_need = steps[t + n4 - 1][3]
                if kx <= k_need or ky < t + n4:
                    continue
                max_ahead_x = max(max_ahead_x, kx - (k_need + 1))
                sg = segs[si]
                n, band, d0 = sg["n"], sg["band"], sg["d0"]
                for j in range(0, n4, 2):
                    tt, ll = t + j, l + j
                    assert steps[tt][:3] == (si, sd, ll) and steps[tt + 1][:3] == (si, sd, ll + 1)
                    sl = tt % Ny
                    assert sl % 2 == 0 and yslot[sl] == tt and yslot[sl + 1] == tt + 1      # adjacent in the dY ring
                    q0 = (3 * ll + sd - 1) % R
                    assert q0 + 12 <= R + MIR                                               # the window never wraps
                    for lr in range(4):
                        for kd in range(3):
                            assert ring[q0 + 3 * lr + kd] == (si, sd - 1 + kd, ll + lr), (tt, lr, kd, ring[q0 + 3 * lr + kd])
                    dp = d0 + sd
                    A = np.zeros((Mm, W))                        # rows line*3Cy + kw*Cy + co
                    for jl in range(2):
                        yl = Y[n, dp, band * LH + 1 + ll + jl][:, :Cy]
                        for kw in range(3):
                            A[jl * 3 * Cy + kw * Cy:jl * 3 * Cy + (kw + 1) * Cy] = yl[1 - kw + 1:1 - kw + 1 + W].T
                    B = np.zeros((12 * Cx, W))                   # columns (window line, kd, ci)
                    for lr in range(4):
                        for kd in range(3):
                            xl = X[n, d0 + sd - 1 + kd, band * LH + ll + lr][:, :Cx]
                            B[(3 * lr + kd) * Cx:(3 * lr + kd + 1) * Cx] = xl[1:1 + W].T
                    acc += A @ B.T
                t += n4
            elif who == 2 and t < nsteps:
                si, sd, l, k_need = steps[t]
                if kx <= k_need or ky <= t:
                    continue
                max_ahead_x = max(max_ahead_x, kx - (k_need + 1))
                sg = segs[si]
                assert yslot[t % Ny] == t
                q0 = (3 * l + sd - 1) % R
                assert q0 + 9 <= R + MIR                         # the window never wraps
                for kh in range(3):
                    for kd in range(3):
                        assert ring[q0 + 3 * kh + kd] == (si, sd - 1 + kd, l + kh), (t, kh, kd, ring[q0 + 3 * kh + kd])
                # ---- operands of this step, as the descriptors address them ----
                n, band, d0 = sg["n"], sg["band"], sg["d0"]
                dp = d0 + sd                                     # padded slice of the dY line (sd = 1 <-> interior slice d0)
                hp = band * LH + 1 + l
                yl = Y[n, dp, hp][:, :Cy]                        # (W+2, Cy): the chunk planes the kernel loads
                A = np.zeros((Mm, W))                            # rows kw*Cy + co, K = X rows 1 .. W
                for kw in range(3):
                    A[kw * Cy:(kw + 1) * Cy] = yl[1 - kw + 1:1 - kw + 1 + W].T      # A_kw[r] = dY[r - kw + 1]
                B = np.zeros((9 * Cx, W))
                for kh in range(3):
                    for kd in range(3):
                        xl = X[n, d0 + sd - 1 + kd, band * LH + l + kh][:, :Cx]      # slice s = sd-1+kd <-> padded d0+s
                        B[(3 * kh + kd) * Cx:(3 * kh + kd + 1) * Cx] = xl[1:1 + W].T
                if 9 * Cx <= 256:
                    acc += A @ B.T
                else:                                            # one MMA per kh: ring slots q0 + 3*kh .. + 2, column block kh
                    for kh in range(3):
                        acc[:, kh * 3 * Cx:(kh + 1) * 3 * Cx] += A @ B[kh * 3 * Cx:(kh + 1) * 3 * Cx].T
                t += 1
    assert max_ahead_x < NB
    if PAIR:
        # the epilogue's two partial blocks: line 0 of a pair owns the columns of window lines 0..2, line 1 those of 1..3
        return acc[:3 * Cy, :9 * Cx] + acc[3 * Cy:6 * Cy, 3 * Cx:12 * Cx]
    return acc


def replay(dy, x, p, seed=0):
    N, _, D, H, W = x.shape
    assert p["units"] == N * p["n_bands"] * D and p["ksteps"] * 16 == W and p["R"] == 3 * (p["LH"] + 2) + 1
    assert p["Lp"] == (W + 2) * 16 and p["smem"] <= 227 * 1024
    nchy, nchx = p["nchy"], p["nchx"]
    Cy, Cx = 8 * nchy, 8 * nchx
    assert Cy <= dy.shape[1] and Cx <= x.shape[1]
    ln = 2 if p["pair"] else 1
    slack = (64 if ln * 3 * Cy <= 64 else 128) // 8 - ln * 3 * nchy  # planes an A operand reads past the last expanded line
    if p["pair"]:
        assert p["LH"] % 2 == 0 and p["Ny"] % 2 == 0 and H % 2 == 0 and p["mirror"] == 11
    assert p["smem_raw_off"] >= (p["R"] + p["mirror"]) * nchx * p["Lp"] and p["smem_y_off"] >= p["smem_raw_off"] and p["NR"] == 0
    assert p["smem"] <= 227 * 1024 - 12 * 1024, "leave shared memory for the co-resident memory-bound kernels"
    assert p["smem_bar_off"] >= p["smem_y_off"] + (p["Ny"] * 3 * nchy + slack) * p["Lp"] and p["LH"] + 2 <= p["ND"]
    Y, X = padded(dy), padded(x)
    rng = random.Random(seed)
    partial = np.stack([replay_cta(p, segments(p, cta, N, D, H), Y, X, W, rng) for cta in range(p["grid"])])
    # wgrad_line_reduce_kernel: dW[co][ci][kd][kh][kw] = sum_cta P[cta][kw*Cy + co][kh*3*Cx + kd*Cx + ci]
    tot = partial.sum(0)
    dW = np.zeros((Cy, Cx, 3, 3, 3))
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                dW[:, :, kd, kh, kw] = tot[kw * Cy:(kw + 1) * Cy, kh * 3 * Cx + kd * Cx:kh * 3 * Cx + kd * Cx + Cx]
    return dW


# (padded channels, real output channels, real input channels): 16 <-> 16, 32 <-> 32 (three MMAs per K step), conv_input
# (4 real input channels: X's upper chunk is skipped, N = 72), conv_output (3 real output channels: dY's upper chunk skipped)
@pytest.mark.parametrize("chans", [(16, 16, 16), (32, 32, 32), (16, 16, 4), (16, 3, 16)])
@pytest.mark.parametrize("shape", [(1, 3, 5, 16), (2, 9, 8, 16), (1, 5, 19, 32), (2, 2, 2, 16), (1, 1, 7, 48),
                                   (1, 3, 6, 16), (1, 4, 12, 32), (2, 3, 10, 16)])
@pytest.mark.parametrize("pair", [True, False])
def test_replay_matches_autograd(shape, chans, pair):
    N, D, H, W = shape
    Cc, ro, ri = chans
    p = line_plan(N, D, H, W, Cc, ro, ri, pair=pair)
    assert (p["nchy"], p["nchx"]) == (1 if ro <= 8 else Cc // 8, 1 if ri <= 8 else Cc // 8)
    # pair mode (two dY lines per MMA, opt-in) applies to 16-channel operands and even H
    assert p["pair"] == int(pair and Cc == 16 and H % 2 == 0)
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, ri, D, H, W, generator=g, dtype=torch.float64)
    w = torch.randn(ro, ri, 3, 3, 3, generator=g, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(N, ro, D, H, W, generator=g, dtype=torch.float64)
    (ref,) = torch.autograd.grad(F.conv3d(x, w, padding=1), w, dy)
    pad = lambda t: F.pad(t, (0, 0, 0, 0, 0, 0, 0, Cc - t.shape[1]))          # channels padded with zeros, as in HBM
    got = replay(pad(dy), pad(x), p, seed=sum(shape))
    np.testing.assert_allclose(got[:ro, :ri], ref.numpy(), rtol=1e-9, atol=1e-9)


def test_schedule_survives_many_interleavings_at_the_benchmark_shape():
    """Config-3 geometry (2 x 128^3): only the ring / barrier bookkeeping is replayed (operands skipped)."""
    for pair in (True, False):
        _schedule_at_benchmark_shape(pair)


def _schedule_at_benchmark_shape(pair):
    p = line_plan(2, 128, 128, 128, pair=pair)
    if pair:
        assert p["pair"] == 1 and p["LH"] % 2 == 0 and 128 % p["LH"] == 0 and p["Ny"] >= 4 and p["Ny"] % 2 == 0
    else:
        assert p["pair"] == 0 and p["LH"] == 8
    assert p["grid"] == 148 and p["Ny"] >= 4 and p["NR"] == 0
    segs = segments(p, 17, 2, 128, 128) + segments(p, 18, 2, 128, 128)      # a CTA-sized run crossing a band boundary

    class NoData:                                                           # index-only stand-in for the activations
        def __getitem__(self, k):
            return np.zeros((130, 16))

    for seed in range(3):
        replay_cta(p, segs, NoData(), NoData(), 128, random.Random(seed))


def test_schedule_at_the_level_1_shape():
    """32 <-> 32 channels at 2 x 64^3 (level 1 of config 3): plan + ring / barrier bookkeeping."""
    p = line_plan(2, 64, 64, 64, 32)
    assert p["nchy"] == 4 and p["nchx"] == 4 and p["grid"] == 148 and p["ksteps"] == 4 and p["LH"] >= 4 and 64 % p["LH"] == 0
    segs = segments(p, 11, 2, 64, 64) + segments(p, 12, 2, 64, 64)

    class NoData:
        def __getitem__(self, k):
            return np.zeros((66, 32))

    for seed in range(3):
        replay_cta(p, segs, NoData(), NoData(), 64, random.Random(seed))


def test_does_not_apply_outside_its_domain():
    for desc in (WgradDesc(0, 1, 8, 8, 24, 16, 16), WgradDesc(0, 1, 8, 8, 32, 64, 64), WgradDesc(0, 1, 8, 8, 32, 32, 16),
                 WgradDesc(1, 1, 8, 8, 32, 16, 16)):
        out = (C.c_int * 32)()
        assert _lib.lib().b200_wgrad_line_plan_debug(C.byref(desc), out, 32) != 0
