"""Inference pipeline around the model, device-resident.

Mirrors the reference procedures without their host round trips:
  * test.py:92-99    pad every spatial dim up to a multiple of 16, centred, zero fill
  * test.py:103-113  z-score per modality with the statistics of the non-zero voxels
  * test.py:115-141  4-flip test-time augmentation (identity, flip D, flip H, flip D+H), averaged
  * test.py:144-159  threshold 0.5, label painting WT->2, TC->1, ET->4 (if ET volume > 32)
  * train.py:145-176 + loader_helper.py:34-97  sliding window: tile 192^3, centre 48^3, border 72;
    tiles are cut from the zero-extended volume and run through the model in batches (GroupNorm
    statistics are per sample, so batching tiles does not change any tile's result); only the
    centre of each tile is written back.
The flips, the averaging and the painting are a handful of elementwise passes over 3-channel
tensors and use torch ops on the GPU; every convolution/normalisation runs in the CUDA library.
The connected-component filter of test.py:160-164 needs skimage and is out of scope.
"""
from __future__ import annotations

import math

import torch


def closest_to_k(n, k=16):
    """loader_helper.py:99-103"""
    return n if n % k == 0 else (n // k + 1) * k


def pad_to_multiple(x, k=16):
    """x: (B,C,D,H,W).  Returns (padded, pad_left, pad_right) like test.py:92-99."""
    old = x.shape[2:]
    new = [closest_to_k(s, k) for s in old]
    left = [(n - o) // 2 for n, o in zip(new, old)]
    right = [(n - o) - l for n, o, l in zip(new, old, left)]
    out = x.new_zeros(x.shape[:2] + tuple(new))
    out[:, :, left[0]:left[0] + old[0], left[1]:left[1] + old[1], left[2]:left[2] + old[2]] = x
    return out, left, right


def unpad(y, left, right):
    d, h, w = y.shape[2:]
    return y[:, :, left[0]:d - right[0], left[1]:h - right[1], left[2]:w - right[2]]


def zscore_nonzero(x):
    """test.py:103-113 per sample and modality: mask = x > 0, mean = sum(x)/n_pos, std from E[x^2]."""
    dims = (2, 3, 4)
    n_pos = (x > 0).sum(dim=dims, keepdim=True).clamp_min(1).to(x.dtype)
    mean = x.sum(dim=dims, keepdim=True) / n_pos
    mean2 = (x * x).sum(dim=dims, keepdim=True) / n_pos
    std = torch.sqrt((mean2 - mean * mean).clamp_min(1e-12))
    return (x - mean) / std


@torch.no_grad()
def predict_tta(model, x):
    """x: (B,4,D,H,W) CUDA fp32, dims multiples of 8.  Average of the 4 flip variants (test.py:115-141)."""
    acc = None
    for dims in ((), (2,), (3,), (2, 3)):
        xi = torch.flip(x, dims).contiguous() if dims else x
        p = model([xi])[0]
        p = torch.flip(p, dims) if dims else p.clone()
        acc = p if acc is None else acc.add_(p)
    return acc / 4.0


def paint_labels(probs, et_min_voxels=32):
    """probs: (B,3,D,H,W) -> uint8 label map (B,D,H,W): WT->2, TC->1, ET->4 (test.py:144-159)."""
    m = probs > 0.5
    out = torch.zeros(probs.shape[:1] + probs.shape[2:], dtype=torch.uint8, device=probs.device)
    out[m[:, 0]] = 2
    out[m[:, 1]] = 1
    for b in range(probs.shape[0]):
        if int(m[b, 2].sum()) > et_min_voxels:
            out[b][m[b, 2]] = 4
    return out


def tile_positions(shape, center):
    """train.py:158: grid = ceil(dim / centre) tiles per axis."""
    grid = [int(math.ceil(s / c)) for s, c in zip(shape, center)]
    return [(i, j, k) for i in range(grid[0]) for j in range(grid[1]) for k in range(grid[2])]


@torch.no_grad()
def predict_tiled(model, x, n_outputs=3, tile=(192, 192, 192), center=(48, 48, 48), border=(72, 72, 72), batch_tiles=2):
    """Sliding-window prediction of ONE volume x: (1,4,D,H,W) CUDA -> (1,n_outputs,D,H,W).
    Same tile geometry as Trainer.predict_tiled (train.py:145-176); input and output stay on the GPU."""
    assert x.shape[0] == 1 and all(t == c + 2 * b for t, c, b in zip(tile, center, border))
    vol = x.shape[2:]
    pos = tile_positions(vol, center)
    grid_max = [max(p[a] for p in pos) + 1 for a in range(3)]
    # zero-extended copy: `border` before, and up to the end of the last tile after (loader_helper.copy)
    ext = [b + g * c + b for b, g, c in zip(border, grid_max, center)]
    xe = x.new_zeros(x.shape[:2] + tuple(ext))
    xe[:, :, border[0]:border[0] + vol[0], border[1]:border[1] + vol[1], border[2]:border[2] + vol[2]] = x
    out = x.new_zeros((1, n_outputs) + tuple(vol))
    for s in range(0, len(pos), batch_tiles):
        chunk = pos[s:s + batch_tiles]
        tiles = torch.cat([xe[:, :, p[0] * center[0]:p[0] * center[0] + tile[0],
                              p[1] * center[1]:p[1] * center[1] + tile[1],
                              p[2] * center[2]:p[2] * center[2] + tile[2]] for p in chunk], dim=0).contiguous()
        pr = model([tiles])[0]
        for b, p in enumerate(chunk):
            lo = [p[a] * center[a] for a in range(3)]
            hi = [min(lo[a] + center[a], vol[a]) for a in range(3)]           # loader_helper.copy_back clamps
            out[0, :, lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = pr[b, :, border[0]:border[0] + hi[0] - lo[0],
                                                                border[1]:border[1] + hi[1] - lo[1],
                                                                border[2]:border[2] + hi[2] - lo[2]]
    return out
