"""Drop-in mirror of the reference's `model.py` hot path (model.py:7-14, 66-117, 308-433).

Same class names, constructor signatures, sub-module and parameter names, so that
`load_state_dict` of a reference checkpoint, `model.apply(weight_init)`, `.cuda()`,
`.train()/.eval()`, pickling via `torch.save`, and `train.Trainer` all work unchanged
(SURVEY.md 8b).  The sub-modules are real `nn.Conv3d` / `nn.GroupNorm` objects used purely as
parameter containers; `UNet.forward` never calls them - it runs the whole network as
hand-written sm_100a kernels through `brats2019_b200.engine.Engine`.

There is no CPU path and no cuDNN path: a CPU tensor, a non-sm_100 device or a missing
libbrats_b200.so raises RuntimeError.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .engine import Engine


class Trilinear(nn.Module):
    """model.py:7-14.  Kept for state-dict/module-tree parity (`upsampling.i.0`); the
    interpolation itself runs inside the engine (b200_upsample2x)."""

    def __init__(self, scale):
        super(Trilinear, self).__init__()
        self.scale = scale

    def forward(self, x):
        raise RuntimeError("brats2019_b200.Trilinear is a structural placeholder; call UNet.forward")


class conv(nn.Module):
    """model.py:66-79: wrapper that adds the `.conv1.conv1.` level to parameter names."""

    def __init__(self, in_channels, out_channels, stride=1, groups=1):
        super(conv, self).__init__()
        if stride != 1 or groups != 1:
            raise RuntimeError("brats2019_b200: only stride=1, groups=1 3x3x3 convs are on the hot path")
        self.conv1 = nn.Conv3d(in_channels=in_channels, out_channels=out_channels, kernel_size=(3, 3, 3),
                               stride=stride, padding=1, bias=False, groups=groups)

    def forward(self, x):
        raise RuntimeError("brats2019_b200.conv is a parameter container; call UNet.forward")


class Residual(nn.Module):
    """model.py:81-117: parameter container with the reference's attribute names."""

    def __init__(self, in_channels, out_channels, stride, downsample=None, conv_groups=1):
        super(Residual, self).__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.downsample = downsample
        self.conv1 = conv(in_channels=in_channels, out_channels=out_channels, stride=stride)
        self.conv2 = conv(in_channels=out_channels, out_channels=out_channels, stride=1)
        self.relu1 = nn.LeakyReLU(1e-2, inplace=True)
        self.relu2 = nn.LeakyReLU(1e-2, inplace=True)
        self.norm1 = nn.GroupNorm(num_groups=8, num_channels=out_channels)
        self.norm2 = nn.GroupNorm(num_groups=8, num_channels=out_channels)

    def forward(self, x):
        raise RuntimeError("brats2019_b200.Residual is a parameter container; call UNet.forward")


class _UNetFunction(torch.autograd.Function):
    """One autograd node for the whole network: forward saves activations inside the engine's
    buffers, backward returns the gradients of the live parameters (dead ones get None, like
    the reference: SURVEY.md 0)."""

    @staticmethod
    def forward(ctx, module, x, names, *params):
        eng = module._engine()
        probs = eng.forward(x, training=True)
        ctx.module = module
        ctx.names = names
        ctx.state = eng.saved_state()       # (plan, generation, probs): which activations this node owns
        return probs

    @staticmethod
    def backward(ctx, gprobs):
        eng = ctx.module._engine()
        factory = ctx.module._grad_store_factory
        grads = eng.backward(gprobs, factory() if factory is not None else None, state=ctx.state)
        ctx.state = None
        return (None, None, None) + tuple(grads.get(n) for n in ctx.names)


class UNet(nn.Module):
    """model.py:308-433 — same constructor, same parameters, `forward([x]) -> [probs]`."""

    def __init__(self, depth, encoder_layers, decoder_layers, number_of_channels, number_of_outputs, block=Residual):
        super(UNet, self).__init__()
        print('UNet {}'.format(number_of_channels))
        if block is not Residual and getattr(block, "__name__", "") != "Residual":
            raise RuntimeError("brats2019_b200: only block=Residual is implemented (main.py:59 default)")
        self.encoder_layers = encoder_layers
        self.decoder_layers = decoder_layers
        self.number_of_channels = number_of_channels
        self.number_of_outputs = number_of_outputs
        self.depth = depth
        self.block = Residual

        self.conv_input = 0
        self.encoder_convs = nn.ModuleList()
        self.upsampling = nn.ModuleList()
        self.decoder_convs = nn.ModuleList()
        self.decoder_convs1x1 = nn.ModuleList()
        self.attention_convs = nn.ModuleList()
        self.upsampling_distance = nn.ModuleList()

        self.conv_input = nn.Conv3d(in_channels=4, out_channels=self.number_of_channels[0], kernel_size=(3, 3, 3),
                                    stride=1, padding=(1, 1, 1), bias=False)
        self.norm_input = nn.GroupNorm(num_groups=8, num_channels=self.number_of_channels[0])
        conv_first_list = []
        for i in range(self.encoder_layers[0]):
            conv_first_list.append(self.block(in_channels=self.number_of_channels[0],
                                              out_channels=self.number_of_channels[0], stride=1))
        self.conv_first = nn.Sequential(*conv_first_list)
        self.conv_output = nn.Conv3d(in_channels=self.number_of_channels[0], out_channels=self.number_of_outputs,
                                     kernel_size=3, stride=1, padding=1, bias=True, groups=1)
        self.softmax = nn.Softmax(dim=1)
        self.sigmoid = nn.Sigmoid()
        self.relu = nn.LeakyReLU(1e-2, inplace=True)

        self.construct_dencoder_convs(depth=depth, number_of_channels=number_of_channels)
        self.construct_encoder_convs(depth=depth, number_of_channels=number_of_channels)
        self.construct_upsampling_convs(depth=depth, number_of_channels=number_of_channels)

        self._grad_store_factory = None     # set by parallel.DistributedUNet (bucketed all-reduce)

    # -- construction helpers, same structure as model.py:358-404 --------------------------
    def _make_encoder_layer(self, in_channels, channels, blocks, stride=1, block=Residual):
        downsample = None
        if stride != 1:
            downsample = nn.Sequential(
                nn.Conv3d(in_channels=in_channels, out_channels=channels, kernel_size=2, stride=stride, bias=False))
        layers = [block(in_channels=channels, out_channels=channels, stride=1, downsample=downsample)]
        for _ in range(1, blocks):
            layers.append(block(in_channels=channels, out_channels=channels, stride=1))
        return nn.Sequential(*layers)

    def construct_encoder_convs(self, depth, number_of_channels):
        for i in range(depth - 1):
            self.encoder_convs.append(self._make_encoder_layer(
                in_channels=number_of_channels[i], channels=number_of_channels[i + 1],
                blocks=self.encoder_layers[i + 1], stride=2, block=self.block))

    def construct_dencoder_convs(self, depth, number_of_channels):
        for i in range(depth):
            conv_list = [self.block(in_channels=number_of_channels[i], out_channels=number_of_channels[i], stride=1)
                         for _ in range(self.decoder_layers[i])]
            self.decoder_convs.append(nn.Sequential(*conv_list))
            self.decoder_convs1x1.append(nn.Conv3d(in_channels=2 * number_of_channels[i],
                                                   out_channels=number_of_channels[i], kernel_size=1, padding=0,
                                                   bias=False))

    def construct_upsampling_convs(self, depth, number_of_channels):
        for i in range(depth - 1):
            self.upsampling.append(nn.Sequential(
                Trilinear(scale=2),
                nn.Conv3d(in_channels=number_of_channels[i + 1], out_channels=number_of_channels[i], kernel_size=1,
                          stride=1, bias=False)))

    # -- engine plumbing -----------------------------------------------------------------------
    def _engine(self):
        eng = self.__dict__.get("_eng")
        if eng is None:
            eng = Engine(self)
            self.__dict__["_eng"] = eng      # not a sub-module, not pickled with state_dict
        return eng

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_eng", None)                  # buffers/plans are rebuilt lazily after unpickling
        d["_grad_store_factory"] = None
        return d

    # Weight edits that do not bump a parameter's version counter (engine.Engine.invalidate_packed): the paths the
    # reference itself uses are covered here - `model.apply(weight_init)` (main.py:60; weight_init.py:22-27 writes
    # through `.data`), `.cuda()` / `.to()` (train.py:54-56) and `load_state_dict` (train.py:333).
    def invalidate_packed_weights(self):
        eng = self.__dict__.get("_eng")
        if eng is not None:
            eng.invalidate_packed()

    def apply(self, fn):
        out = super(UNet, self).apply(fn)
        self.invalidate_packed_weights()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super(UNet, self)._apply(fn, *args, **kwargs)
        self.invalidate_packed_weights()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super(UNet, self).load_state_dict(*args, **kwargs)
        self.invalidate_packed_weights()
        return out

    def dead_parameter_names(self):
        """Parameters the reference builds but never uses (model.py:379-395 vs :420)."""
        i = self.depth - 1
        names = ["decoder_convs1x1.%d.weight" % i]
        names += [n for n, _ in self.named_parameters() if n.startswith("decoder_convs.%d." % i)]
        return names

    def live_parameters(self):
        dead = set(self.dead_parameter_names())
        return [(n, p) for n, p in self.named_parameters() if n not in dead]

    def forward(self, x, return_logits=False):
        """x: list whose first element is a CUDA fp32 (B,4,D,H,W) tensor (model.py:407-410).
        Returns [probabilities (B, n_out, D, H, W) fp32] (model.py:431-433)."""
        inp = x[0]
        if not (torch.is_tensor(inp) and inp.is_cuda):
            raise RuntimeError("brats2019_b200.UNet runs on a B200 (sm_100a) only; got a %s tensor - there is no CPU "
                               "or cuDNN fallback" % (inp.device if torch.is_tensor(inp) else type(inp)))
        if inp.dim() != 5:
            raise RuntimeError("expected (B,4,D,H,W), got %s" % (tuple(inp.shape),))
        with torch.cuda.device(inp.device):
            if return_logits:
                with torch.no_grad():
                    probs, logits = self._engine().forward(inp, want_logits=True)
                return [probs], logits
            live = self.live_parameters()
            if torch.is_grad_enabled() and any(p.requires_grad for _, p in live):
                names = tuple(n for n, _ in live)
                probs = _UNetFunction.apply(self, inp, names, *[p for _, p in live])
            else:
                with torch.no_grad():
                    probs = self._engine().forward(inp)
        return [probs]
