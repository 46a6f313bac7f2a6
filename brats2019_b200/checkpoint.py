"""Checkpoint compatibility with the reference trainer.

The reference saves `{'state': train.TrainingState, 'model': nn.DataParallel(model.UNet)}` with
`torch.save` of the whole objects (train.py:320-324), so its files are pickles that name the classes
`model.UNet`, `model.Residual`, `model.conv`, `model.Trilinear` and `train.TrainingState`, and every
parameter key carries DataParallel's `module.` prefix (export_onnx_group_norm.py:27-31 strips it).

`load_reference_checkpoint` reads such a file WITHOUT the reference sources on `sys.path`: an unpickler
maps the reference's class names onto this package's drop-in classes (same constructor-built sub-module
tree and parameter names, so the pickled module state restores directly) and onto a plain
`TrainingState` record.  `save_reference_layout` writes the same two-key layout back, the model wrapped
so that its `state_dict()` keys start with `module.` - a reference script that does
`s['model'].state_dict()` (train.py:331-333, export_onnx_group_norm.py:27-31) reads it unchanged.
"""
from __future__ import annotations

import pickle
from collections import OrderedDict

import torch
import torch.nn as nn

from . import model as _model


class TrainingState(object):
    """train.py:14-24: the bookkeeping record the reference trainer pickles next to the model."""

    def __init__(self):
        self.epoch = 0
        self.train_metric = dict()
        self.val_metric = dict()
        self.global_step = 0
        self.best_val = 0
        self.optimizer_state = None
        self.cuda = True


class ModulePrefix(nn.Module):
    """Holds the network as `.module`, like nn.DataParallel (main.py:61), without its scatter/gather: its
    state_dict keys are `module.<name>`, which is the key layout of every reference checkpoint."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return self.module(x)


_CLASS_MAP = {
    ("model", "UNet"): _model.UNet,
    ("model", "Residual"): _model.Residual,
    ("model", "conv"): _model.conv,
    ("model", "Trilinear"): _model.Trilinear,
    ("train", "TrainingState"): TrainingState,
    ("torch.nn.parallel.data_parallel", "DataParallel"): ModulePrefix,
}


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, mod, name):
        hit = _CLASS_MAP.get((mod, name))
        if hit is not None:
            return hit
        if mod in ("model", "train", "loss", "metrics", "weight_init"):
            raise pickle.UnpicklingError("reference class %s.%s has no drop-in in brats2019_b200" % (mod, name))
        return super().find_class(mod, name)


class _RefPickleModule:
    """`pickle_module` for torch.load: the stock pickle with the class-remapping unpickler."""
    __name__ = "brats2019_b200.checkpoint._RefPickleModule"
    Unpickler = _RefUnpickler
    load = staticmethod(pickle.load)
    loads = staticmethod(pickle.loads)
    dump = staticmethod(pickle.dump)
    dumps = staticmethod(pickle.dumps)
    HIGHEST_PROTOCOL = pickle.HIGHEST_PROTOCOL
    UnpicklingError = pickle.UnpicklingError
    PicklingError = pickle.PicklingError


def strip_module_prefix(state_dict):
    """export_onnx_group_norm.py:28-31: drop DataParallel's `module.` from every key."""
    out = OrderedDict()
    for k, v in state_dict.items():
        out[k[7:] if k.startswith("module.") else k] = v
    return out


def load_reference_checkpoint(path, map_location="cpu"):
    """Returns (UNet, TrainingState) from a file written by the reference's Trainer._save
    (train.py:320-324) or by `save_reference_layout`.  The UNet is this package's drop-in class with the
    checkpoint's weights; move it to the GPU with `.cuda()` as the reference callers do."""
    s = torch.load(path, map_location=map_location, pickle_module=_RefPickleModule, weights_only=False)
    if not (isinstance(s, dict) and "model" in s):
        raise RuntimeError("not a reference checkpoint (expected a dict with 'model' and 'state'): %s" % path)
    net = s["model"]
    net = net.module if hasattr(net, "module") else net
    if not isinstance(net, _model.UNet):
        raise RuntimeError("checkpoint model is %s, expected the residual UNet" % type(net).__name__)
    # Rebuild from the constructor arguments so that attributes added by this package (engine plumbing) exist,
    # then take the weights by name - the same route export_onnx_group_norm.py:26-32 takes.
    fresh = _model.UNet(depth=net.depth, encoder_layers=list(net.encoder_layers), decoder_layers=list(net.decoder_layers),
                        number_of_channels=list(net.number_of_channels), number_of_outputs=net.number_of_outputs)
    fresh.load_state_dict(strip_module_prefix(net.state_dict()))
    return fresh, s.get("state")


def load_state_dict_into(model, path, map_location="cpu"):
    """Trainer._load with an existing model (train.py:331-333): copy the checkpoint's weights into `model`
    (a UNet, or a wrapper exposing `.module`)."""
    src, state = load_reference_checkpoint(path, map_location)
    target = model.module if hasattr(model, "module") else model
    target.load_state_dict(src.state_dict())
    return state


def save_reference_layout(path, model, state=None):
    """Trainer._save (train.py:320-324): {'state': ..., 'model': <module whose keys start with 'module.'>}."""
    net = model.module if hasattr(model, "module") else model
    torch.save({"state": state if state is not None else TrainingState(), "model": ModulePrefix(net)}, path)
