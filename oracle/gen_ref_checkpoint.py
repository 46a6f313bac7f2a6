"""Write a checkpoint exactly the way the reference trainer does (run in the build container only).

    python oracle/gen_ref_checkpoint.py        # writes tests/golden/ref_checkpoint_tiny.pth (+ .npz of a forward)

Imports the REAL /root/reference/model.py, builds a small residual UNet (depth 2, channels [16, 32] so the
file stays ~0.5 MB), initialises it with the reference's weight_init.py, wraps it in nn.DataParallel
(main.py:61) and saves {'state': TrainingState, 'model': DataParallel} with torch.save of the whole objects
(train.py:320-324).  `train.py` itself needs tensorboardX, so a module named `train` holding a TrainingState
class with the fields of train.py:14-24 stands in for it: the pickle only records the class path and __dict__.
Also stores the reference's CPU forward on a seeded input, so the test can check that the weights arrived."""
import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, "/root/reference")
import model as ref_model  # noqa: E402
import weight_init  # noqa: E402

train = types.ModuleType("train")


class TrainingState(object):
    def __init__(self):
        self.epoch = 0
        self.train_metric = dict()
        self.val_metric = dict()
        self.global_step = 0
        self.best_val = 0
        self.optimizer_state = None
        self.cuda = True


TrainingState.__module__ = "train"
train.TrainingState = TrainingState
sys.modules["train"] = train

torch.manual_seed(1337)
net = ref_model.UNet(depth=2, encoder_layers=[1, 2], decoder_layers=[1, 1], number_of_channels=[16, 32], number_of_outputs=3)
net.apply(weight_init.weight_init)
state = TrainingState()
state.epoch, state.global_step, state.best_val = 7, 7000, np.array([0.9, 0.8, 0.7])
out = os.path.join(REPO, "tests", "golden", "ref_checkpoint_tiny.pth")
torch.save({"state": state, "model": torch.nn.DataParallel(module=net, device_ids=[0])}, out)
g = torch.Generator().manual_seed(3)
x = torch.randn(1, 4, 16, 16, 16, generator=g)
with torch.no_grad():
    p = net([x])[0]
np.savez_compressed(os.path.join(REPO, "tests", "golden", "ref_checkpoint_tiny_forward.npz"), x=x.numpy(), probs=p.numpy())
print(out, os.path.getsize(out))
