"""Thin Python wrappers over the C ABI: torch tensors in, raw device pointers out.

PyTorch is used here for device memory and streams only.  Every function launches
hand-written sm_100a kernels from libbrats_b200.so on `torch.cuda.current_stream()`.
"""
from __future__ import annotations

import ctypes as C

import os

import torch

from . import _lib
from ._lib import ConvDesc, WgradDesc, check

MODE_K3, MODE_K1 = 0, 1
EPI_BF16, EPI_SIGMOID, EPI_D2S = 0, 1, 2
W_FWD, W_DGRAD, W_FWD_S2D, W_DGRAD_S2D = 0, 1, 2, 3
G_K3, G_K1, G_S2D = 0, 1, 2
GN_EPS = 1e-5

# kernels of ours enqueued so far (bench.py reports the count inside its timed region)
LAUNCHES = [0]


def _count(n):
    LAUNCHES[0] += n


# Work ledger (bench.py `roofline_all`): when LEDGER is a list, every wrapper appends
# (kernel-name prefix as CUPTI reports it, "tensor" | "hbm", algorithmic FLOPs or bytes of that launch) - the
# algorithmic work of SURVEY.md 8(d): 2*M*N*K per conv with padding taps counted as full taps, and
# (elements read + written) * element size at the kernel's logical interface for the memory-bound kernels.
LEDGER = None


def _ledger(name, bound, work):
    if LEDGER is not None:
        LEDGER.append((name, bound, float(work)))


def _act_bytes(a):
    return 2.0 * a.N * a.D * a.H * a.W * a.C


def _conv_kernel_name(desc):
    """Which kernel b200_conv_run dispatches this descriptor to (same planners, no launch)."""
    L = _lib.lib()
    buf = (C.c_int * 64)()
    if desc.mode == MODE_K3 and L.b200_band_plan_debug(C.byref(desc), buf, 64) == 0:
        return "conv_band_kernel"
    if desc.mode == MODE_K3 and L.b200_march_plan_debug(C.byref(desc), buf, 64) == 0:
        return "conv_march_kernel"
    return "conv_gemm_kernel<%d," % desc.mode


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return None if t is None else t.data_ptr()


def pad16(c):
    return (c + 15) // 16 * 16


# ------------------------------------------------------------------------------------------
# activation buffers: chunk-planar, zero halos, zero guard rows (csrc/common.cuh):
#   t[C/8][guard + N*(D+2)*(H+2)*(W+2) + guard][8]  bf16
# ------------------------------------------------------------------------------------------
class Act:
    """One activation tensor in the library's HBM layout (a torch buffer + its geometry)."""
    __slots__ = ("t", "N", "D", "H", "W", "C", "G", "PR")

    def __init__(self, N, D, H, W, Cc, device):
        assert Cc % 8 == 0
        L = _lib.lib()
        self.N, self.D, self.H, self.W, self.C = N, D, H, W, Cc
        self.G = L.b200_act_guard_rows(D, H, W)
        self.PR = L.b200_act_plane_rows(N, D, H, W)
        self.t = torch.zeros((Cc // 8, self.PR, 8), dtype=torch.bfloat16, device=device)

    def data_ptr(self):
        return self.t.data_ptr()

    def chunks(self, c0, c1):
        """View of the 8-channel chunk planes [c0, c1) as an activation tensor of 8*(c1-c0) channels (the
        layout is chunk-planar, so a channel range that starts on a chunk is itself a valid tensor)."""
        v = Act.__new__(Act)
        v.N, v.D, v.H, v.W, v.C, v.G, v.PR = self.N, self.D, self.H, self.W, 8 * (c1 - c0), self.G, self.PR
        v.t = self.t[c0:c1]
        return v

    @property
    def device(self):
        return self.t.device

    def padded(self):
        """View (C/8, N, D+2, H+2, W+2, 8) of the rows between the guards."""
        rows = self.N * (self.D + 2) * (self.H + 2) * (self.W + 2)
        return self.t[:, self.G:self.G + rows].view(self.C // 8, self.N, self.D + 2, self.H + 2, self.W + 2, 8)

    def interior(self):
        return self.padded()[:, :, 1:-1, 1:-1, 1:-1, :]

    def nbytes(self):
        return self.t.numel() * 2


def act_zeros(N, D, H, W, Cc, device):
    return Act(N, D, H, W, Cc, device)


def act_dims(a):
    return a.N, a.D, a.H, a.W, a.C


def act_from_ncdhw(x, Cpad=None):
    """Reference-side helper (tests): fp32 NCDHW -> act, done with torch ops."""
    N, Cc, D, H, W = x.shape
    Cpad = Cpad or pad16(Cc)
    a = Act(N, D, H, W, Cpad, x.device)
    xp = torch.zeros((N, Cpad, D, H, W), dtype=torch.bfloat16, device=x.device)
    xp[:, :Cc] = x.to(torch.bfloat16)
    a.interior().copy_(xp.view(N, Cpad // 8, 8, D, H, W).permute(1, 0, 3, 4, 5, 2))
    return a


def act_to_ncdhw(a, Cc=None):
    """Reference-side helper (tests): act interior -> fp32 NCDHW."""
    Cc = Cc or a.C
    x = a.interior().permute(1, 0, 5, 2, 3, 4).reshape(a.N, a.C, a.D, a.H, a.W)
    return x[:, :Cc].float().contiguous()


def act_outside_absmax(a):
    """Reference-side helper (tests): max |value| over halo and guard elements (must stay 0)."""
    c = a.t.clone()
    rows = a.N * (a.D + 2) * (a.H + 2) * (a.W + 2)
    c[:, a.G:a.G + rows].view(a.C // 8, a.N, a.D + 2, a.H + 2, a.W + 2, 8)[:, :, 1:-1, 1:-1, 1:-1, :] = 0
    return c.float().abs().max()


F32, BF16, U8 = 0, 1, 2          # include/brats_b200.h: B200_F32 / B200_BF16 / B200_U8
_DT = {torch.float32: F32, torch.bfloat16: BF16, torch.uint8: U8, torch.bool: U8}


def pack_input(x, Cpad=16, out=None):
    N, Cc, D, H, W = x.shape
    assert x.dtype in (torch.float32, torch.bfloat16) and x.is_contiguous() and x.is_cuda
    if out is None:
        out = act_zeros(N, D, H, W, Cpad, x.device)
    assert Cc <= 8 and out.C == Cpad
    check(_lib.lib().b200_pack_input_t(_p(x), _DT[x.dtype], _p(out), N, D, H, W, Cc, Cpad, _stream()), "b200_pack_input")
    _count(1)
    _ledger("pack_input", "hbm", x.numel() * x.element_size() + N * D * H * W * 16)
    return out


# ------------------------------------------------------------------------------------------
# convolution
# ------------------------------------------------------------------------------------------
def conv_desc(mode, N, D, H, W, Cin_a, Cout, Cin_b=0, epi=EPI_BF16):
    return ConvDesc(mode, epi, N, D, H, W, Cin_a, Cin_b, Cout)


def conv_ctas(desc):
    n = _lib.lib().b200_conv_ctas(C.byref(desc))
    if n < 0:
        check(1, "b200_conv_ctas")
    return n


def conv_pack_weight(desc, kind, w, ci_off=0, K_real=None, N_real=None, out=None):
    """w: fp32 (Cout_w, Cin_w, *kernel) cuda tensor -> packed bf16 operand images (uint8 buffer)."""
    L = _lib.lib()
    nbytes = L.b200_conv_packed_weight_bytes(C.byref(desc))
    if nbytes == 0:
        check(1, "b200_conv_packed_weight_bytes")
    w = w.detach()
    assert w.dtype == torch.float32 and w.is_cuda
    w = w.contiguous()
    Cout_w, Cin_w = w.shape[0], w.shape[1]
    taps = 1
    for s in w.shape[2:]:
        taps *= s
    if K_real is None or N_real is None:
        if kind == W_FWD:
            k, n = Cin_w - ci_off, Cout_w
        elif kind == W_DGRAD:
            k, n = Cout_w, Cin_w - ci_off
        elif kind == W_FWD_S2D:
            k, n = 8 * Cin_w, Cout_w
        else:
            k, n = Cout_w, 8 * Cin_w
        K_real = k if K_real is None else K_real
        N_real = n if N_real is None else N_real
    if out is None:
        out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    assert out.numel() >= nbytes
    check(L.b200_conv_pack_weight(C.byref(desc), kind, _p(w), Cout_w, Cin_w, taps, ci_off, K_real, N_real, _p(out),
                                  _stream()), "b200_conv_pack_weight")
    _count(1)
    return out


class PackTable:
    """All packed weight images of one model; `run()` refreshes every image with ONE kernel launch
    (b200_pack_table_run) from a device-resident job table that is rebuilt only when a job is added or
    a parameter's storage moves."""

    def __init__(self):
        self.jobs = []          # [desc, kind, w, ci_off, K_real, N_real, buf]
        self.table = None
        self.ptrs = None
        self.total_blocks = 0

    def add(self, desc, kind, w, ci_off, K_real, N_real):
        buf = conv_pack_weight(desc, kind, w, ci_off=ci_off, K_real=K_real, N_real=N_real)
        self.jobs.append([desc, kind, w, ci_off, K_real, N_real, buf])
        self.table = None
        if not torch.cuda.is_current_stream_capturing():
            self._build()          # keep the device table current outside graph capture (first step only)
        return len(self.jobs) - 1, buf

    def repack_one(self, idx):
        desc, kind, w, ci_off, K_real, N_real, buf = self.jobs[idx]
        conv_pack_weight(desc, kind, w, ci_off=ci_off, K_real=K_real, N_real=N_real, out=buf)

    def _build(self):
        L = _lib.lib()
        n = len(self.jobs)
        arr = (_lib.PackJob * n)()
        for i, (desc, kind, w, ci_off, K_real, N_real, buf) in enumerate(self.jobs):
            assert w.dtype == torch.float32 and w.is_cuda and w.is_contiguous()
            taps = 1
            for sdim in w.shape[2:]:
                taps *= sdim
            arr[i].desc = desc
            arr[i].kind, arr[i].Cout_w, arr[i].Cin_w, arr[i].taps_w = kind, w.shape[0], w.shape[1], taps
            arr[i].ci_off, arr[i].K_real, arr[i].N_real = ci_off, K_real, N_real
            arr[i].w, arr[i].packed = w.data_ptr(), buf.data_ptr()
        nbytes = n * L.b200_pack_table_entry_bytes()
        host = (C.c_uint8 * nbytes)()
        total = C.c_int(0)
        check(L.b200_pack_table_build(arr, n, host, nbytes, C.byref(total)), "b200_pack_table_build")
        dev = self.jobs[0][2].device
        if torch.cuda.is_current_stream_capturing():
            # pinned staging buffer (kept alive): the copy is legal while a CUDA graph is being captured
            self._host = torch.frombuffer(bytearray(host), dtype=torch.uint8).pin_memory()
            self.table = self._host.to(dev, non_blocking=True)
        else:
            # A few KB, built once per model: a plain synchronous copy.  (Until round 2c this path pinned the buffer too;
            # in a long process - the whole GPU test suite - torch's pinned-memory allocator then raised a one-shot
            # "CUDA error: invalid argument" here while polling the events of earlier, already freed staging buffers.)
            self._host = None
            self.table = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(dev)
        self.total_blocks = total.value
        self.ptrs = [j[2].data_ptr() for j in self.jobs]

    def run(self):
        if not self.jobs:
            return
        if self.table is None or self.ptrs != [j[2].data_ptr() for j in self.jobs]:
            self._build()
        check(_lib.lib().b200_pack_table_run(_p(self.table), len(self.jobs), self.total_blocks, _stream()),
              "b200_pack_table_run")
        _count(1)
        _ledger("pack_weights_batched_kernel", "hbm", sum(j[2].numel() * 4 + j[6].numel() for j in self.jobs))


def conv_supports_gnbwd(desc):
    return _lib.lib().b200_conv_supports_gnbwd(C.byref(desc)) != 0


def conv_run(desc, src_a, packed, out=None, src_b=None, residual=None, lrelu=False, stats=None, bias=None,
             probs=None, logits=None, n_out_real=0, gnb_x=None, gnb_coef=None):
    """gnb_x / gnb_coef: fold the backward sums of the GroupNorm that consumes `out` as its dy into the epilogue
    (b200_conv_run_gnbwd); `stats` then receives those sums."""
    check(_lib.lib().b200_conv_run_gnbwd(C.byref(desc), _p(src_a), _p(src_b), _p(packed), _p(out), _p(residual),
                                         1 if lrelu else 0, _p(stats), _p(bias), _p(probs), _p(logits), n_out_real,
                                         _p(gnb_x), _p(gnb_coef), _stream()), "b200_conv_run")
    _count(1)
    if LEDGER is not None:
        taps = 27 if desc.mode == MODE_K3 else 1
        _ledger(_conv_kernel_name(desc), "tensor",
                2.0 * desc.N * desc.D * desc.H * desc.W * (desc.Cin_a + desc.Cin_b) * desc.Cout * taps)
    return out


def conv_run_gn(desc, src, packed, out, stats, mean, rstd, ticket):
    """3x3x3 conv + GroupNorm statistics in one launch (b200_conv_run_gn): the last CTA reduces `stats` to mean / rstd.
    ticket: zero-initialised int32 tensor (one word is used and left zero)."""
    check(_lib.lib().b200_conv_run_gn(C.byref(desc), _p(src), _p(packed), _p(out), _p(stats), _p(mean), _p(rstd),
                                      _p(ticket), GN_EPS, _stream()), "b200_conv_run_gn")
    _count(1)
    if LEDGER is not None:
        _ledger(_conv_kernel_name(desc), "tensor", 2.0 * desc.N * desc.D * desc.H * desc.W * desc.Cin_a * desc.Cout * 27)
    return out


def wgrad_desc(mode, N, D, H, W, Cout, Cin):
    return WgradDesc(mode, N, D, H, W, Cout, Cin)


def wgrad_workspace(desc, device):
    nbytes = _lib.lib().b200_wgrad_workspace_bytes(C.byref(desc))
    if nbytes == 0:
        check(1, "b200_wgrad_workspace_bytes")
    return torch.empty(nbytes // 4, dtype=torch.float32, device=device)


def wgrad_run(desc, dy, x, grad, kind, ci_off=0, accumulate=False, workspace=None):
    """grad: fp32 (Cout_w, Cin_w, *kernel) contiguous; written (or accumulated) in place."""
    if workspace is None:
        workspace = wgrad_workspace(desc, dy.device)
    assert grad.dtype == torch.float32 and grad.is_contiguous()
    taps = 1
    for s in grad.shape[2:]:
        taps *= s
    check(_lib.lib().b200_wgrad_run(C.byref(desc), _p(dy), _p(x), _p(workspace), _p(grad), kind, grad.shape[0],
                                    grad.shape[1], taps, ci_off, 1 if accumulate else 0, _stream()),
          "b200_wgrad_run")
    _count(2)
    if LEDGER is not None:
        import os
        buf = (C.c_int * 64)()
        line = (kind == G_K3 and os.environ.get("B200_NO_WGRAD_LINE", "0") in ("", "0") and
                os.environ.get("B200_WGRAD_MARCH", "0") in ("", "0") and
                (desc.Cout != 32 or os.environ.get("B200_WGRAD_LINE32", "0") not in ("", "0")) and
                _lib.lib().b200_wgrad_line_plan_debug(C.byref(desc), buf, 64) == 0)
        t = 27 if desc.mode == 0 else 1
        _ledger("wgrad_line_kernel" if line else "wgrad_gemm_kernel", "tensor",
                2.0 * desc.N * desc.D * desc.H * desc.W * desc.Cout * desc.Cin * t)
        _ledger("wgrad_line_reduce_kernel" if line else "wgrad_reduce_kernel", "hbm", grad.numel() * 4)
    return grad


# ------------------------------------------------------------------------------------------
# GroupNorm / activation / residual
# ------------------------------------------------------------------------------------------
def gn_finalize(stats, ctas, N, Cc, D, H, W, mean, rstd, gamma=None, beta=None, coef=None, lrelu=True):
    """coef (with gamma, beta): also write the [N][3][C] table the backward fold reads (b200_gn_finalize_coef)."""
    if coef is not None:
        check(_lib.lib().b200_gn_finalize_coef(_p(stats), ctas, N, Cc, D, H, W, GN_EPS, _p(gamma), _p(beta),
                                               1 if lrelu else 0, _p(mean), _p(rstd), _p(coef), _stream()),
              "b200_gn_finalize_coef")
    else:
        check(_lib.lib().b200_gn_finalize(_p(stats), ctas, N, Cc, D, H, W, GN_EPS, _p(mean), _p(rstd), _stream()),
              "b200_gn_finalize")
    _count(1)
    _ledger("gn_finalize_kernel", "hbm", ctas * N * 16 * 4)


def gn_apply(x, mean, rstd, gamma, beta, out, residual=None, lrelu=True):
    N, D, H, W, Cc = act_dims(x)
    check(_lib.lib().b200_gn_apply(_p(x), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(residual), _p(out), N, D, H, W,
                                   Cc, 1 if lrelu else 0, _stream()), "b200_gn_apply")
    _count(1)
    _ledger("gn_apply_kernel", "hbm", (3 if residual is not None else 2) * _act_bytes(x))
    return out


def gn_backward_workspace(N, Cc, device):
    n = _lib.lib().b200_gn_backward_workspace_floats(N, Cc)
    return torch.zeros(n, dtype=torch.float32, device=device)     # its leading ticket words must start at zero


def gn_backward(x, dy, mean, rstd, gamma, beta, dx, dgamma, dbeta, workspace, lrelu=True):
    N, D, H, W, Cc = act_dims(x)
    check(_lib.lib().b200_gn_backward(_p(x), _p(dy), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(dx), _p(dgamma),
                                      _p(dbeta), _p(workspace), N, D, H, W, Cc, 1 if lrelu else 0, _stream()),
          "b200_gn_backward")
    form = _lib.lib().b200_gn_backward_form(N, D, H, W, Cc)
    if form == 2:
        _count(1)
        _ledger("gn_bwd_creg_kernel", "hbm", 3 * _act_bytes(x))
        return dx
    if form == 1:
        _count(1)
        _ledger("gn_bwd_cluster_kernel", "hbm", 3 * _act_bytes(x))
        return dx
    fused = Cc <= 128 and os.environ.get("B200_GN_BWD_FUSED_FIN", "0") not in ("", "0")
    _count(2 if fused else 3)
    _ledger("gn_bwd_reduce2_kernel", "hbm", 2 * _act_bytes(x))
    if not fused:
        _ledger("gn_bwd_finalize2_kernel", "hbm", 0)
    _ledger("gn_bwd_apply2_kernel", "hbm", 3 * _act_bytes(x))
    return dx


def gn_backward_folded(x, dy, mean, rstd, gamma, beta, gpart, ctas, dx, dgamma, dbeta, workspace, lrelu=True):
    """GroupNorm backward whose group sums `gpart` were left by the conv that produced dy (conv_run(gnb_x=...))."""
    N, D, H, W, Cc = act_dims(x)
    check(_lib.lib().b200_gn_backward_folded(_p(x), _p(dy), _p(mean), _p(rstd), _p(gamma), _p(beta), _p(gpart), ctas,
                                             _p(dx), _p(dgamma), _p(dbeta), _p(workspace), N, D, H, W, Cc,
                                             1 if lrelu else 0, _stream()), "b200_gn_backward_folded")
    _count(3)
    _ledger("gn_bwd_fold_finalize_kernel", "hbm", 0)
    _ledger("gn_bwd_apply2_kernel", "hbm", 3 * _act_bytes(x))
    _ledger("gn_bwd_finalize2_kernel", "hbm", 0)
    return dx


# ------------------------------------------------------------------------------------------
# resampling / layout
# ------------------------------------------------------------------------------------------
def upsample2x(coarse, fine, lrelu=True):
    N, D, H, W, Cc = act_dims(coarse)
    check(_lib.lib().b200_upsample2x(_p(coarse), _p(fine), N, D, H, W, Cc, 1 if lrelu else 0, _stream()),
          "b200_upsample2x")
    _count(1)
    _ledger("upsample2x_fwd3_kernel", "hbm", 9 * _act_bytes(coarse))
    return fine


_UPS_WS = {}


def free_workspaces():
    """Drop the cached scratch buffers of this module (Engine.free_plans)."""
    _UPS_WS.clear()


def upsample2x_backward(dfine, fine_out, dcoarse, lrelu=True, workspace=None):
    N, D, H, W, Cc = act_dims(dcoarse)
    L = _lib.lib()
    if workspace is None:
        # one cached scratch buffer per (device, size): allocation-free after the first call (CUDA graphs)
        nbytes = L.b200_upsample2x_backward_workspace_bytes(N, D, H, W, Cc)
        key = (dcoarse.device, nbytes)
        workspace = _UPS_WS.get(key)
        if workspace is None:
            workspace = torch.empty(nbytes, dtype=torch.uint8, device=dcoarse.device)
            _UPS_WS[key] = workspace
    check(L.b200_upsample2x_backward(_p(dfine), _p(fine_out), _p(dcoarse), _p(workspace), N, D, H, W, Cc,
                                     1 if lrelu else 0, _stream()), "b200_upsample2x_backward")
    _count(2)
    _ledger("upsample2x_bwd_w3_kernel", "hbm", (8 + 8 + 4) * _act_bytes(dcoarse))
    _ledger("upsample2x_bwd_dh3_kernel", "hbm", (4 + 1) * _act_bytes(dcoarse))
    return dcoarse


def space_to_depth(fine, coarse):
    N, D, H, W, C8 = act_dims(coarse)
    check(_lib.lib().b200_space_to_depth(_p(fine), _p(coarse), N, D, H, W, C8 // 8, _stream()), "b200_space_to_depth")
    _count(1)
    _ledger("s2d_kernel", "hbm", 2 * _act_bytes(coarse))
    return coarse


def depth_to_space(coarse, fine, residual=None):
    N, D, H, W, C8 = act_dims(coarse)
    check(_lib.lib().b200_depth_to_space(_p(coarse), _p(residual), _p(fine), N, D, H, W, C8 // 8, _stream()),
          "b200_depth_to_space")
    _count(1)
    _ledger("d2s_kernel", "hbm", (3 if residual is not None else 2) * _act_bytes(coarse))
    return fine


def add(a, b, out):
    N, D, H, W, Cc = act_dims(a)
    check(_lib.lib().b200_add(_p(a), _p(b), _p(out), N, D, H, W, Cc, _stream()), "b200_add")
    _count(1)
    return out


# ------------------------------------------------------------------------------------------
# sigmoid backward / Dice
# ------------------------------------------------------------------------------------------
def sigmoid_backward(grad_probs, probs, dlogit_act, dbias, workspace=None):
    N, Cr, D, H, W = probs.shape
    Cpad = dlogit_act.C
    if workspace is None:
        workspace = torch.empty(_lib.lib().b200_sigmoid_backward_workspace_floats(N, D, H), dtype=torch.float32,
                                device=probs.device)
    check(_lib.lib().b200_sigmoid_backward(_p(grad_probs), _p(probs), _p(dlogit_act), _p(dbias), _p(workspace), N, D,
                                           H, W, Cr, Cpad, _stream()), "b200_sigmoid_backward")
    _count(2)
    _ledger("sigmoid_bwd_pack", "hbm", 2 * probs.numel() * 4 + N * D * H * W * 16)
    return dlogit_act


def dice_sums(probs, target, sums=None, workspace=None):
    B, Cc = probs.shape[:2]
    S = probs[0, 0].numel()
    if sums is None:
        sums = torch.zeros(8, dtype=torch.float32, device=probs.device)
    if workspace is None:
        workspace = torch.empty(_lib.lib().b200_dice_workspace_floats(B, Cc), dtype=torch.float32, device=probs.device)
    check(_lib.lib().b200_dice_sums_t(_p(probs), _p(target), _DT[target.dtype], _p(sums), _p(workspace), B, Cc, S,
                                      _stream()), "b200_dice_sums")
    _count(2)
    _ledger("dice_partial_kernel", "hbm", probs.numel() * (4 + target.element_size()))
    return sums


def dice_loss(sums, Cc, priority, loss=None):
    if loss is None:
        loss = torch.empty(1, dtype=torch.float32, device=sums.device)
    check(_lib.lib().b200_dice_loss(_p(sums), Cc, float(priority), _p(loss), _stream()), "b200_dice_loss")
    _count(1)
    return loss


def dice_backward(probs, target, sums, grad_out, priority, grad_probs=None):
    B, Cc = probs.shape[:2]
    S = probs[0, 0].numel()
    if grad_probs is None:
        grad_probs = torch.empty_like(probs)
    check(_lib.lib().b200_dice_backward_t(_p(probs), _p(target), _DT[target.dtype], _p(sums), _p(grad_out),
                                          float(priority), _p(grad_probs), B, Cc, S, _stream()), "b200_dice_backward")
    _count(1)
    _ledger("dice_bwd_kernel", "hbm", probs.numel() * (8 + target.element_size()))
    return grad_probs


def bce_sum(probs, target, bg_weight, workspace=None):
    L = _lib.lib()
    s = torch.empty(1, dtype=torch.float32, device=probs.device)
    if workspace is None:
        workspace = torch.empty(L.b200_bce_workspace_floats(), dtype=torch.float32, device=probs.device)
    check(L.b200_bce_sum_t(_p(probs), _p(target), _DT[target.dtype], float(bg_weight), _p(s), _p(workspace),
                           probs.numel(), _stream()), "b200_bce_sum")
    _count(2)
    return s


def bce_loss(s, global_numel):
    loss = torch.empty(1, dtype=torch.float32, device=s.device)
    check(_lib.lib().b200_bce_loss(_p(s), float(global_numel), _p(loss), _stream()), "b200_bce_loss")
    _count(1)
    return loss


def bce_backward(probs, target, grad_out, bg_weight, global_numel):
    gp = torch.empty_like(probs)
    check(_lib.lib().b200_bce_backward_t(_p(probs), _p(target), _DT[target.dtype], _p(grad_out), float(bg_weight),
                                         float(global_numel), _p(gp), probs.numel(), _stream()), "b200_bce_backward")
    _count(1)
    return gp
