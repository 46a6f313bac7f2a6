"""CPU oracle for the distilled residual 3D U-Net hot path (TEST INFRASTRUCTURE ONLY).

This file is a from-scratch, functional restatement (plain torch fp32 ops on CPU or any
device) of the reference's algorithm for the hot path.  It is the *checker*: only `tests/`,
`__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference` legs of `bench.py` may
import it.  The product path (`brats2019_b200`) never does.

Parity pin: the reference ships no tests or golden vectors, so this oracle is
pinned against outputs of the reference itself, generated in the build container by
`oracle/gen_golden.py` (imports /root/reference/model.py + loss.py + weight_init.py) and
committed under `tests/golden/`.  `tests/test_oracle.py` checks oracle == golden.

Reference lines followed:
  model.py:7-14     Trilinear            -> `trilinear_x2`
  model.py:66-79    conv (k3 p1 no bias) -> `F.conv3d(..., padding=1)`
  model.py:81-117   Residual             -> `residual_block`
  model.py:308-433  UNet ctor + forward  -> `param_shapes`, `unet_forward`
  loss.py:98-122    Dice_loss_joint      -> `dice_loss_joint`
  loss.py:64-79     BCE_Loss             -> `bce_loss`
  weight_init.py:22-27                   -> `init_params`
  test.py:144, metrics.py:108-133        -> `threshold_masks`, `dice_metric`
  test.py:115-141                        -> `tta_predict`
  train.py:145-176, loader_helper.py:34-97 -> `tile_get_indices`, `tile_copy`, `tile_copy_back`, `predict_tiled`
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch
import torch.nn.functional as F

# main.py:56-59 — the only configuration the reference ships.
DEFAULT_CFG = dict(
    depth=4,
    encoder_layers=[1, 2, 2, 4],
    decoder_layers=[1, 1, 1, 1],
    number_of_channels=[16, 32, 64, 128],
    number_of_outputs=3,
)
GN_GROUPS = 8          # model.py:95-96, 338
GN_EPS = 1e-5          # nn.GroupNorm default
LRELU_SLOPE = 1e-2     # model.py:93-94, 352


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
def _residual_names(prefix, c):
    # model.py:81-96: conv wrapper adds the `.conv1.conv1.` level (model.py:72)
    return [
        (prefix + "conv1.conv1.weight", (c, c, 3, 3, 3)),
        (prefix + "conv2.conv1.weight", (c, c, 3, 3, 3)),
        (prefix + "norm1.weight", (c,)),
        (prefix + "norm1.bias", (c,)),
        (prefix + "norm2.weight", (c,)),
        (prefix + "norm2.bias", (c,)),
    ]


def param_shapes(cfg=DEFAULT_CFG):
    """Ordered (name, shape) list in the reference's `state_dict()` order.

    Order follows module registration order in model.py:309-356: encoder_convs, upsampling,
    decoder_convs, decoder_convs1x1 lists are registered (empty) first, then conv_input,
    norm_input, conv_first, conv_output; the lists are filled afterwards, which does not
    change their position.
    """
    ch = cfg["number_of_channels"]
    depth = cfg["depth"]
    enc, dec = cfg["encoder_layers"], cfg["decoder_layers"]
    out = []
    # encoder_convs (model.py:374-377, 358-372)
    for i in range(depth - 1):
        c = ch[i + 1]
        for j in range(enc[i + 1]):
            p = "encoder_convs.%d.%d." % (i, j)
            names = _residual_names(p, c)
            if j == 0:
                names = [(p + "downsample.0.weight", (c, ch[i], 2, 2, 2))] + names
            out += names
    # upsampling (model.py:397-404)
    for i in range(depth - 1):
        out.append(("upsampling.%d.1.weight" % i, (ch[i], ch[i + 1], 1, 1, 1)))
    # decoder_convs / decoder_convs1x1 (model.py:379-395): `depth` of each, last one dead
    for i in range(depth):
        for j in range(dec[i]):
            out += _residual_names("decoder_convs.%d.%d." % (i, j), ch[i])
    for i in range(depth):
        out.append(("decoder_convs1x1.%d.weight" % i, (ch[i], 2 * ch[i], 1, 1, 1)))
    out.append(("conv_input.weight", (ch[0], 4, 3, 3, 3)))
    out.append(("norm_input.weight", (ch[0],)))
    out.append(("norm_input.bias", (ch[0],)))
    for j in range(enc[0]):
        out += _residual_names("conv_first.%d." % j, ch[0])
    out.append(("conv_output.weight", (cfg["number_of_outputs"], ch[0], 3, 3, 3)))
    out.append(("conv_output.bias", (cfg["number_of_outputs"],)))
    return out


def dead_param_names(cfg=DEFAULT_CFG):
    """Parameters built but never used by forward (model.py:420 iterates depth-1 only)."""
    i = cfg["depth"] - 1
    names = ["decoder_convs1x1.%d.weight" % i]
    for j in range(cfg["decoder_layers"][i]):
        names += [n for n, _ in _residual_names("decoder_convs.%d.%d." % (i, j), 0)]
    return names


def init_params(seed, cfg=DEFAULT_CFG):
    """Deterministic parameters with the *distribution* of weight_init.py:22-27.

    Conv3d weight ~ kaiming_normal_(a=1e-2, 'leaky_relu', fan_in), Conv3d bias ~ N(0,1),
    GroupNorm weight=1, bias=0.  The draw order is our own (name order of `param_shapes`),
    so values differ from `model.apply(weight_init)` under the same seed; parity tests load
    the same state dict into both sides instead of relying on RNG order.
    """
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    for name, shape in param_shapes(cfg):
        if len(shape) == 5:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            gain = math.sqrt(2.0 / (1 + LRELU_SLOPE ** 2))
            sd[name] = torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))
        elif name.endswith("conv_output.bias"):
            sd[name] = torch.randn(shape, generator=g)
        elif name.endswith(".weight"):
            sd[name] = torch.ones(shape)
        else:
            sd[name] = torch.zeros(shape)
    return sd


# ----------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------
def trilinear_x2(x):
    # model.py:13 — F.interpolate(scale_factor=2, mode='trilinear'), align_corners default False
    return F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=False)


def trilinear_x2_explicit(x):
    """Same op written as the separable fixed-tap stencil:
    out[2k] = .25 in[k-1] + .75 in[k]; out[2k+1] = .75 in[k] + .25 in[k+1], edge-clamped."""
    for dim in (2, 3, 4):
        n = x.shape[dim]
        idx = torch.arange(n)
        lo = x.index_select(dim, (idx - 1).clamp(min=0))
        hi = x.index_select(dim, (idx + 1).clamp(max=n - 1))
        even = 0.25 * lo + 0.75 * x
        odd = 0.75 * x + 0.25 * hi
        x = torch.stack([even, odd], dim=dim + 1).flatten(dim, dim + 1)
    return x


def _gn_lrelu(x, w, b):
    return F.leaky_relu(F.group_norm(x, GN_GROUPS, w, b, GN_EPS), LRELU_SLOPE)


def residual_block(p, prefix, x):
    # model.py:99-117
    if prefix + "downsample.0.weight" in p:
        x = F.conv3d(x, p[prefix + "downsample.0.weight"], stride=2)
    out = F.conv3d(x, p[prefix + "conv1.conv1.weight"], padding=1)
    out = _gn_lrelu(out, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"])
    out = F.conv3d(out, p[prefix + "conv2.conv1.weight"], padding=1)
    out = _gn_lrelu(out, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"])
    return x + out


def unet_logits(p, x, cfg=DEFAULT_CFG):
    """model.py:407-429 up to (not including) the sigmoid.  x: (B,4,D,H,W) fp32."""
    depth = cfg["depth"]
    enc, dec = cfg["encoder_layers"], cfg["decoder_layers"]
    h = F.conv3d(x, p["conv_input.weight"], padding=1)                       # model.py:412
    h = F.group_norm(h, GN_GROUPS, p["norm_input.weight"], p["norm_input.bias"], GN_EPS)  # :413, no act
    for j in range(enc[0]):                                                   # :414
        h = residual_block(p, "conv_first.%d." % j, h)
    skips = []
    for i in range(depth - 1):                                                # :416-418
        skips.append(h)
        for j in range(enc[i + 1]):
            h = residual_block(p, "encoder_convs.%d.%d." % (i, j), h)
    for i in reversed(range(depth - 1)):                                      # :420-426
        h = F.conv3d(trilinear_x2(h), p["upsampling.%d.1.weight" % i])        # :421
        h = F.leaky_relu(h, LRELU_SLOPE)                                      # :422
        h = torch.cat([skips[i], h], dim=1)                                   # :424
        h = F.conv3d(h, p["decoder_convs1x1.%d.weight" % i])                  # :425
        for j in range(dec[i]):                                               # :426
            h = residual_block(p, "decoder_convs.%d.%d." % (i, j), h)
    return F.conv3d(h, p["conv_output.weight"], p["conv_output.bias"], padding=1)  # :429


def unet_forward(p, x_list, cfg=DEFAULT_CFG):
    """List-in / list-out like model.py:407-433: returns [sigmoid(logits)]."""
    return [torch.sigmoid(unet_logits(p, x_list[0], cfg))]


# ----------------------------------------------------------------------------------------
# losses / metrics
# ----------------------------------------------------------------------------------------
def dice_loss_joint(x_list, y_list, index=0, priority=1):
    # loss.py:105-122
    pred, gt = x_list[index], y_list[index]
    assert pred.shape == gt.shape
    n, c = pred.shape[:2]
    pred = pred.reshape(n, c, -1)
    gt = gt.reshape(n, c, -1)
    inter = (pred * gt).sum(dim=(0, 2)) + 1e-6
    union = (pred ** 2 + gt).sum(dim=(0, 2)) + 2e-6
    return priority * (1.0 - torch.mean(2.0 * inter / union))


def bce_loss(x_list, y_list, index=0, bg_weight=1.0):
    # loss.py:70-79
    pred, gt = x_list[index], y_list[index]
    loss = gt * torch.log(pred + 1e-6) + bg_weight * (1.0 - gt) * torch.log((1.0 + 1e-6) - pred)
    return -torch.mean(loss)


def dice_loss_grad_closed_form(pred, gt):
    """dL/dpred of dice_loss_joint in closed form, used to check the
    CUDA backward without autograd: -(2/C) (g U_c - 2 p I_c) / U_c^2."""
    n, c = pred.shape[:2]
    dims = [0] + list(range(2, pred.dim()))
    inter = (pred * gt).sum(dim=dims, keepdim=True) + 1e-6
    union = (pred ** 2 + gt).sum(dim=dims, keepdim=True) + 2e-6
    return -(2.0 / c) * (gt * union - 2.0 * pred * inter) / union ** 2


def threshold_masks(probs):
    # test.py:144 — per-channel threshold, no argmax
    return probs > 0.5


def dice_metric(pred_mask, gt_mask):
    """metrics.py:108-133 (class Dice): per-sample, per-channel thresholded Dice with the
    both-empty case counted as 1."""
    n, c = pred_mask.shape[:2]
    p = pred_mask.reshape(n, c, -1).float()
    g = gt_mask.reshape(n, c, -1).float()
    inter = (p * g).sum(-1)
    denom = p.sum(-1) + g.sum(-1)
    return torch.where(denom > 0, 2.0 * inter / denom.clamp(min=1.0), torch.ones_like(denom))


# ----------------------------------------------------------------------------------------
# training step (forward + loss + autograd backward) — the oracle for gradients
# ----------------------------------------------------------------------------------------
def train_step(p, x, target, with_bce=False, cfg=DEFAULT_CFG):
    """Returns (loss, probs, {name: grad}) for the live parameters (train.py:201-210)."""
    dead = set(dead_param_names(cfg))
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items() if k not in dead)
    probs = unet_forward(leaves, [x], cfg)
    loss = dice_loss_joint(probs, [target])
    if with_bce:                                   # main.py:126-128 + train.py:203-205
        loss = (loss + bce_loss(probs, [target], bg_weight=1e-2)) / 2
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), probs[0].detach(), OrderedDict(zip(leaves.keys(), grads))


# ----------------------------------------------------------------------------------------
# inference procedures around the model
# ----------------------------------------------------------------------------------------
def tta_predict(p, x, cfg=DEFAULT_CFG):
    """test.py:115-141: identity, flip D, flip H, flip D+H; un-flip; average."""
    outs = []
    for dims in ((), (2,), (3,), (2, 3)):
        xi = torch.flip(x, dims) if dims else x
        o = unet_forward(p, [xi], cfg)[0]
        outs.append(torch.flip(o, dims) if dims else o)
    return sum(outs) / len(outs)


def tile_get_indices(position, center_shape, border):
    # loader_helper.py:34-39
    index = [p * c for p, c in zip(position, center_shape)]
    index_min = [i - b for i, b in zip(index, border)]
    index_max = [i + c + b for i, c, b in zip(index, center_shape, border)]
    return index_min, index_max


def tile_copy(data, tile_shape, index_min, index_max):
    # loader_helper.py:42-58: window [index_min, index_max) with zero fill outside the volume
    ret = torch.zeros(tuple(data.shape[:2]) + tuple(tile_shape), dtype=data.dtype, device=data.device)
    shape = data.shape[2:]
    cmin = [max(i, 0) for i in index_min]
    cmax = [min(i, s) for i, s in zip(index_max, shape)]
    dmin = [c - i for c, i in zip(cmin, index_min)]
    dmax = [t - (i - c) for t, c, i in zip(tile_shape, cmax, index_max)]
    ret[:, :, dmin[0]:dmax[0], dmin[1]:dmax[1], dmin[2]:dmax[2]] = data[:, :, cmin[0]:cmax[0], cmin[1]:cmax[1], cmin[2]:cmax[2]]
    return ret


def tile_copy_back(data, tile, center_shape, index_min, index_max, border):
    # loader_helper.py:82-97: only the centre of the tile is written, clamped to the volume
    shape = data.shape[2:]
    cen_min = [i + b for i, b in zip(index_min, border)]
    cen_max = [i - b for i, b in zip(index_max, border)]
    cmin = [max(i, 0) for i in cen_min]
    cmax = [min(i, s) for i, s in zip(cen_max, shape)]
    dmin = [b + c - i for b, c, i in zip(border, cmin, cen_min)]
    dmax = [b + t - (i - c) for b, t, c, i in zip(border, center_shape, cmax, cen_max)]
    data[:, :, cmin[0]:cmax[0], cmin[1]:cmax[1], cmin[2]:cmax[2]] = tile[:, :, dmin[0]:dmax[0], dmin[1]:dmax[1], dmin[2]:dmax[2]]


def predict_tiled(fn, x, output_shape, tile_shape=(192, 192, 192), center_shape=(48, 48, 48), border=(72, 72, 72)):
    """train.py:145-176 with `fn(tile) -> tile-shaped output` standing for self.model([tile])[0]."""
    out = torch.zeros(output_shape, dtype=x.dtype, device=x.device)
    grid = [int(math.ceil(j / i)) for i, j in zip(center_shape, x.shape[2:])]
    for i in range(grid[0]):
        for j in range(grid[1]):
            for k in range(grid[2]):
                imin, imax = tile_get_indices((i, j, k), center_shape, border)
                t = tile_copy(x, tile_shape, imin, imax)
                tile_copy_back(out, fn(t), center_shape, imin, imax, border)
    return out


# ----------------------------------------------------------------------------------------
# bf16-storage emulation of the CUDA path ("separate algorithmic bugs from rounding")
# ----------------------------------------------------------------------------------------
# The product stores every activation tensor, every activation-gradient tensor and the packed weight
# operands in bf16 and accumulates in fp32; the reference is fp32 throughout, so a comparison against the
# fp32 oracle can never be tighter than the bf16 noise floor (~1 % of max |logit|, ~0.3 % of voxels
# flipping side of the 0.5 threshold).  The functions below restate the SAME reference algorithm
# (model.py:99-117, 407-431; loss.py) in fp32 torch ops but round to bf16 exactly where the kernels store:
#   forward   packed input (b200_pack_input), packed weights, every conv output, every GroupNorm-apply
#             (+LeakyReLU, +residual) output, the 1x1-before-upsample output and the upsample output
#             (the engine runs Conv1x1(trilinear(x)) as trilinear(Conv1x1(x)): exact in real arithmetic);
#             GroupNorm statistics come from the conv's fp32 accumulators, not from the rounded tensor;
#   backward  the gradient w.r.t. every stored activation (outputs of gn_bwd_apply2, of the data-gradient
#             convs, of depth_to_space, both halves of the cat conv's data gradient, the W-reduced
#             intermediate of the trilinear adjoint, dlogit); weight / affine gradients stay fp32.
# With `rounding(False)` every rounding is the identity and the functions must reproduce the fp32
# oracle / golden vectors to float precision: that pins the restructured algebra (hand-written GroupNorm
# and trilinear backward, commuted 1x1) to the reference before any CUDA result is compared with it.
_ROUND = [True]


class rounding:
    """Context manager: `with rounding(False):` turns the bf16 roundings of the emulation off."""

    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev, _ROUND[0] = _ROUND[0], self.on

    def __exit__(self, *a):
        _ROUND[0] = self.prev


def _bf(x):
    return x.to(torch.bfloat16).to(torch.float32) if _ROUND[0] else x


class _Stored(torch.autograd.Function):
    """A tensor that lives in a bf16 buffer, and so does its gradient."""

    @staticmethod
    def forward(ctx, x):
        return _bf(x)

    @staticmethod
    def backward(ctx, g):
        return _bf(g)


class _PackedWeight(torch.autograd.Function):
    """bf16 operand image of an fp32 parameter; the weight gradient is accumulated and kept in fp32."""

    @staticmethod
    def forward(ctx, w):
        return _bf(w)

    @staticmethod
    def backward(ctx, g):
        return g


class _GradStored(torch.autograd.Function):
    """fp32 forward value whose gradient is written to a bf16 buffer (dlogit, dskip, ds2d)."""

    @staticmethod
    def forward(ctx, x):
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return _bf(g)


def _per_channel(v, C, ndim):
    """(N, 8) group values -> (N, C, 1, 1, 1) per-channel values."""
    N = v.shape[0]
    return v.repeat_interleave(C // GN_GROUPS, dim=1).reshape((N, C) + (1,) * (ndim - 2))


def gn_stats_emulated(c32):
    """gn_finalize (csrc/elementwise.cuh): per (sample, group) mean and 1/sqrt(var + eps) of the conv's fp32
    accumulators, biased variance as E[x^2] - mean^2 in double (aten::native_group_norm statistics)."""
    N = c32.shape[0]
    xg = c32.double().reshape(N, GN_GROUPS, -1)
    m = xg.mean(-1)
    var = ((xg * xg).mean(-1) - m * m).clamp_min(0.0)
    return m.float(), (1.0 / torch.sqrt(var + GN_EPS)).float()


def gn_apply_emulated(c, mean, rstd, gamma, beta, lrelu):
    """gn_apply (csrc/elementwise.cuh) before the optional residual add and the bf16 store."""
    C = c.shape[1]
    g = gamma.reshape((1, C) + (1,) * (c.dim() - 2))
    b = beta.reshape((1, C) + (1,) * (c.dim() - 2))
    scale = _per_channel(rstd, C, c.dim()) * g
    z = c * scale + (b - _per_channel(mean, C, c.dim()) * scale)
    return F.leaky_relu(z, LRELU_SLOPE) if lrelu else z


def gn_backward_emulated(c, mean, rstd, gamma, beta, dy, lrelu):
    """gn_bwd_reduce2 / finalize2 / apply2 (csrc/elementwise2.cuh): closed-form backward of
    y = lrelu?(GroupNorm(c)) with the bf16-stored conv output `c`.  Returns (dx fp32 before the bf16 store,
    dgamma, dbeta)."""
    N, C = c.shape[:2]
    gs = C // GN_GROUPS
    sp = tuple(range(2, c.dim()))
    g = gamma.reshape((1, C) + (1,) * (c.dim() - 2))
    b = beta.reshape((1, C) + (1,) * (c.dim() - 2))
    r = _per_channel(rstd, C, c.dim())
    mu = _per_channel(mean, C, c.dim())
    bb = -mu * r
    p1 = r * g
    z = c * p1 + (bb * g + b)
    dz = dy * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, LRELU_SLOPE)) if lrelu else dy
    xh = c * r + bb
    S1 = dz.double().sum(sp)                      # (N, C)
    S2 = (dz * xh).double().sum(sp)
    count = float(gs * c[0, 0].numel())
    gd = g.reshape(1, C).double()
    A = ((S1 * gd).reshape(N, GN_GROUPS, gs).sum(-1) / count).float()
    B = ((S2 * gd).reshape(N, GN_GROUPS, gs).sum(-1) / count).float()
    Ac, Bc = _per_channel(A, C, c.dim()), _per_channel(B, C, c.dim())
    dx = dz * p1 + (c * (-r * r * Bc) + (-r * (Ac + bb * Bc)))
    return dx, S2.sum(0).float(), S1.sum(0).float()


class _GroupNormEmulated(torch.autograd.Function):
    """gn_finalize + gn_apply forward, gn_bwd_reduce2/finalize2/apply2 backward (csrc/elementwise*.cuh):
    statistics from the fp32 conv accumulators, normalisation of the bf16-stored conv output, the
    closed-form GroupNorm(+LeakyReLU) backward with the bf16-stored input, output gradient in bf16."""

    @staticmethod
    def forward(ctx, c32, gamma, beta, lrelu):
        mean, rstd = gn_stats_emulated(c32)
        c = _bf(c32)
        ctx.save_for_backward(c, mean, rstd, gamma, beta)
        ctx.lrelu = lrelu
        return gn_apply_emulated(c, mean, rstd, gamma, beta, lrelu)

    @staticmethod
    def backward(ctx, dy):
        c, mean, rstd, gamma, beta = ctx.saved_tensors
        dx, dgamma, dbeta = gn_backward_emulated(c, mean, rstd, gamma, beta, dy, ctx.lrelu)
        return _bf(dx), dgamma, dbeta, None


def _up1d(x, dim):
    n = x.shape[dim]
    idx = torch.arange(n, device=x.device)
    lo = x.index_select(dim, (idx - 1).clamp(min=0))
    hi = x.index_select(dim, (idx + 1).clamp(max=n - 1))
    even = 0.25 * lo + 0.75 * x
    odd = 0.75 * x + 0.25 * hi
    return torch.stack([even, odd], dim=dim + 1).flatten(dim, dim + 1)


def _up1d_adjoint(g, dim):
    """Adjoint of `_up1d`: out[k] = .25 g[2k-1] + .75 g[2k] + .75 g[2k+1] + .25 g[2k+2], fine indices clamped to
    the volume (a clamped duplicate adds its weight, exactly as the forward clamp does)."""
    n2 = g.shape[dim]
    k = torch.arange(n2 // 2, device=g.device)
    a = g.index_select(dim, (2 * k - 1).clamp(min=0))
    b = g.index_select(dim, 2 * k)
    c = g.index_select(dim, 2 * k + 1)
    d = g.index_select(dim, (2 * k + 2).clamp(max=n2 - 1))
    return 0.25 * (a + d) + 0.75 * (b + c)


def upsample_lrelu_emulated(x):
    """upsample2x_fwd3 before the bf16 store: trilinear x2 (W, then H, then D) + LeakyReLU."""
    y = x
    for dim in (4, 3, 2):
        y = _up1d(y, dim)
    return F.leaky_relu(y, LRELU_SLOPE)


def upsample_lrelu_backward_emulated(g, pos):
    """upsample2x_bwd_w3 + upsample2x_bwd_dh3: mask by the sign of the stored output, reduce along W, store that
    intermediate in bf16, reduce along H and D.  Returns the fp32 value before the final bf16 store."""
    g = g * torch.where(pos, torch.ones_like(g), torch.full_like(g, LRELU_SLOPE))
    t = _bf(_up1d_adjoint(g, 4))
    return _up1d_adjoint(_up1d_adjoint(t, 3), 2)


class _UpsampleLReluEmulated(torch.autograd.Function):
    """upsample2x_fwd3 / upsample2x_bwd_w3 + upsample2x_bwd_dh3 (csrc/elementwise4.cuh): trilinear x2
    (model.py:7-14) followed by LeakyReLU (model.py:422); the adjoint reduces along W first and stores that
    intermediate in bf16."""

    @staticmethod
    def forward(ctx, x):
        y = x
        for dim in (4, 3, 2):
            y = _up1d(y, dim)
        ctx.save_for_backward(y > 0)
        return F.leaky_relu(y, LRELU_SLOPE)

    @staticmethod
    def backward(ctx, g):
        (pos,) = ctx.saved_tensors
        return upsample_lrelu_backward_emulated(g, pos)


def _conv_e(x, w, **kw):
    return F.conv3d(x, _PackedWeight.apply(w), **kw)


def _residual_block_emulated(p, prefix, x, trace=None):
    """model.py:99-117 on a stored tensor `x` (after the optional downsample)."""
    c1 = _conv_e(x, p[prefix + "conv1.conv1.weight"], padding=1)
    a1 = _Stored.apply(_GroupNormEmulated.apply(c1, p[prefix + "norm1.weight"], p[prefix + "norm1.bias"], True))
    c2 = _conv_e(a1, p[prefix + "conv2.conv1.weight"], padding=1)
    out = _Stored.apply(x + _GroupNormEmulated.apply(c2, p[prefix + "norm2.weight"], p[prefix + "norm2.bias"], True))
    if trace is not None:       # stored tensors under the engine's buffer names (brats2019_b200/engine.py)
        trace[prefix + "c1"], trace[prefix + "a1"], trace[prefix + "c2"], trace[prefix + "out"] = _bf(c1), a1, _bf(c2), out
    return out


def unet_logits_bf16_emulated(p, x, cfg=DEFAULT_CFG, trace=None):
    """`unet_logits` with the storage roundings of the CUDA path (see the section header).  `trace`: optional dict
    that receives every stored intermediate under the engine's buffer name (layer-by-layer diagnosis)."""
    tr = trace if trace is not None else {}
    depth = cfg["depth"]
    enc, dec = cfg["encoder_layers"], cfg["decoder_layers"]
    h = _conv_e(_bf(x), p["conv_input.weight"], padding=1)                                   # model.py:412
    tr["in.c"] = _bf(h)
    h = _Stored.apply(_GroupNormEmulated.apply(h, p["norm_input.weight"], p["norm_input.bias"], False))   # :413
    tr["in.a"] = h
    for j in range(enc[0]):
        h = _residual_block_emulated(p, "conv_first.%d." % j, h, trace)
    skips = []
    for i in range(depth - 1):
        skips.append(h)
        for j in range(enc[i + 1]):
            prefix = "encoder_convs.%d.%d." % (i, j)
            if j == 0:                                                                        # model.py:101-102
                # (the data gradient of this conv is NOT stored on its own: the depth-to-space epilogue adds it to the skip
                # gradient in fp32 and the sum is rounded where the next stage reads it)
                h = _Stored.apply(_conv_e(h, p[prefix + "downsample.0.weight"], stride=2))
                tr["enc%d.down" % i] = h
            h = _residual_block_emulated(p, prefix, h, trace)
    for i in reversed(range(depth - 1)):
        ulo = _Stored.apply(_conv_e(h, p["upsampling.%d.1.weight" % i]))                      # :421, commuted
        up = _Stored.apply(_UpsampleLReluEmulated.apply(ulo))                                 # :421-422
        cat = torch.cat([_GradStored.apply(skips[i]), up], dim=1)                             # :424
        h = _Stored.apply(_conv_e(cat, p["decoder_convs1x1.%d.weight" % i]))                  # :425
        tr["dec%d.ulo" % i], tr["dec%d.up" % i], tr["dec%d.cc" % i] = ulo, up, h
        for j in range(dec[i]):
            h = _residual_block_emulated(p, "decoder_convs.%d.%d." % (i, j), h, trace)
    logits = F.conv3d(h, _PackedWeight.apply(p["conv_output.weight"]), p["conv_output.bias"], padding=1)
    return _GradStored.apply(logits)


def train_step_bf16_emulated(p, x, target, with_bce=False, cfg=DEFAULT_CFG):
    """`train_step` on the emulated path: (loss, probs, logits, {name: grad})."""
    dead = set(dead_param_names(cfg))
    leaves = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in p.items() if k not in dead)
    logits = unet_logits_bf16_emulated(leaves, x, cfg)
    probs = [torch.sigmoid(logits)]
    loss = dice_loss_joint(probs, [target])
    if with_bce:
        loss = (loss + bce_loss(probs, [target], bg_weight=1e-2)) / 2
    grads = torch.autograd.grad(loss, list(leaves.values()))
    return loss.detach(), probs[0].detach(), logits.detach(), OrderedDict(zip(leaves.keys(), grads))
