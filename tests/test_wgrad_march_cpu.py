"""CPU replay of the marching weight-gradient kernel's addressing (brats2019_b200/csrc/wgrad_march.cuh) against
torch's conv3d weight gradient: the planner's numbers (b200_wgrad_march_plan_debug) drive a numpy emulation of
what the CTAs do - unit ranges, segments, the 3-slot dY ring filled in refill order, the accumulator set chosen by
the ring slot of the centre slice, the (slot, chunk) / (line, chunk) block order of the two MN-major operands, the
K ranges over the interior voxels of a line, the kw start shifts - and of the reduce kernel's un-rotation.
No GPU: what the hardware does with the descriptors is covered by tests/gpu_opcheck.py (wgrad group)."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from brats2019_b200 import _lib
from brats2019_b200._lib import WgradDesc

FIELDS = ("BH n_bands units ksteps BHs nsub Wp y_plane_bytes y_sub_bytes y_slot_bytes x_blk_bytes x_stage_bytes "
          "smem_y_off smem_stg_off smem_x_off smem grid").split()


def march_plan(N, D, H, W):
    d = WgradDesc(0, N, D, H, W, 16, 16)
    out = (C.c_int * 32)()
    assert _lib.lib().b200_wgrad_march_plan_debug(C.byref(d), out, 32) == 0, _lib.lib().b200_last_error()
    return {k: out[i] for i, k in enumerate(FIELDS)}


def padded(x):
    """(N,C,D,H,W) -> (N, D+2, H+2, W+2, C) with the zero halo the activation layout keeps in HBM."""
    N, Cc, D, H, W = x.shape
    a = np.zeros((N, D + 2, H + 2, W + 2, Cc), np.float64)
    a[:, 1:-1, 1:-1, 1:-1] = x.permute(0, 2, 3, 4, 1).numpy()
    return a


def replay(dy, x, p):
    N, _, D, H, W = x.shape
    BH, nb, units, grid, ksteps = p["BH"], p["n_bands"], p["units"], p["grid"], p["ksteps"]
    assert units == N * nb * D and ksteps * 16 == W and p["BHs"] * p["nsub"] == BH
    Y, X = padded(dy), padded(x)
    Hp = H + 2
    partial = np.zeros((grid, 9, 64, 48))
    steps_done = 0
    for cta in range(grid):
        u, u_end = units * cta // grid, units * (cta + 1) // grid
        j = 0                                              # dY refills consumed so far (ring slot = refill % 3)
        ring = [None, None, None]
        while u < u_end:
            d0 = u % D
            band = (u // D) % nb
            n = u // D // nb
            length = min(D - d0, u_end - u)
            d1 = d0 + length - 1
            u += length
            nl = min(BH, H - band * BH)

            def y_lines(dyp):                              # the band's BH interior lines of padded slice dyp (may run past H)
                out = np.zeros((BH, W + 2, 16))
                for l in range(BH):
                    hp = band * BH + 1 + l
                    if hp < Hp:
                        out[l] = Y[n, dyp, hp]
                return out
            # refills of this segment: padded dY slices d0 .. d1+2, in order, slot = refill counter % 3
            seg_refills = list(range(d0, d1 + 3))
            loaded = 0

            def refill_until(count):
                nonlocal loaded
                while loaded < count:
                    ring[(j_seg + loaded) % 3] = y_lines(seg_refills[loaded])
                    loaded += 1
            j_seg = j
            for q, dxp in enumerate(range(d0 + 1, d1 + 2)):
                refill_until(q + 3)                        # slices dxp-1, dxp, dxp+1 = refills j, j+1, j+2
                rho = (j + 1) % 3
                xs = np.zeros((BH + 2, W + 2, 16))         # X stage: padded lines band*BH .. +BH+1 of slice dxp
                for ln in range(BH + 2):
                    hp = band * BH + ln
                    if hp < Hp:
                        xs[ln] = X[n, dxp, hp]
                for l in range(nl):
                    for t in range(3):                     # kw tap = accumulator
                        for ks in range(ksteps):
                            rows_y = slice(1 + 16 * ks, 17 + 16 * ks)          # dY rows wp = 1 + 16 ks ..
                            rows_x = slice(16 * ks + t, 16 * ks + t + 16)      # X rows shifted by kw - 1
                            A = np.concatenate([ring[s][l, rows_y] for s in range(3)], axis=1)       # [16 k][48 m = s*16+co]
                            Bm = np.concatenate([xs[l + f, rows_x] for f in range(3)], axis=1)       # [16 k][48 n = kh*16+ci]
                            partial[cta, rho * 3 + t, :48] += A.T @ Bm
                steps_done += 1
                j += 1
            j += 2
    assert steps_done == units
    # reduce kernel: dW[kd][kh][kw][co][ci] = sum_cta sum_rho P[cta][rho*3+kw][s*16+co][kh*16+ci], s = (rho + 4 - kd) % 3
    P = partial.sum(axis=0)
    g = np.zeros((16, 16, 3, 3, 3))
    for kd in range(3):
        for kh in range(3):
            for kw in range(3):
                for rho in range(3):
                    s = (rho + 4 - kd) % 3
                    g[:, :, kd, kh, kw] += P[rho * 3 + kw, s * 16:s * 16 + 16, kh * 16:kh * 16 + 16]
    return g


@pytest.mark.parametrize("shape", [(1, 3, 4, 16), (2, 5, 7, 32), (1, 6, 9, 16), (3, 1, 2, 16)])
def test_wgrad_march_replay_matches_conv3d_weight_gradient(shape):
    N, D, H, W = shape
    p = march_plan(N, D, H, W)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(N, 16, D, H, W, generator=g, dtype=torch.float64)
    dy = torch.randn(N, 16, D, H, W, generator=g, dtype=torch.float64)
    w = torch.zeros(16, 16, 3, 3, 3, dtype=torch.float64, requires_grad=True)
    F.conv3d(x, w, padding=1).backward(dy)
    got = replay(dy, x, p)
    np.testing.assert_allclose(got, w.grad.numpy(), rtol=1e-9, atol=1e-9)


def test_wgrad_march_plan_for_the_bench_shape():
    p = march_plan(2, 128, 128, 128)
    assert p["BH"] == 4 and p["nsub"] == 2 and p["grid"] == 148 and p["units"] == 2 * 32 * 128
    assert p["smem"] <= 227 * 1024
    # the 64-row A operand reads one slot past each half-band ring: that memory must be inside the allocation
    assert p["smem_stg_off"] == 3 * p["y_slot_bytes"] and p["smem_x_off"] == 5 * p["y_slot_bytes"]
    d = WgradDesc(0, 2, 16, 16, 24, 16, 16)               # W % 16 != 0: the linear-row kernel keeps the job
    out = (C.c_int * 32)()
    assert _lib.lib().b200_wgrad_march_plan_debug(C.byref(d), out, 32) != 0
