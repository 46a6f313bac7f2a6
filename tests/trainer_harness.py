"""The reference's training / inference DRIVER for the drop-in tests (SURVEY.md 4 "driver parity", 8b "who calls it").

`load()` returns the reference's own `train.Trainer` (and `metrics` module) when the reference checkout is present
(BRATS_REFERENCE_DIR, default /root/reference) - imported unmodified, with the three absent third-party modules
stubbed from tests/ref_stubs/ - and otherwise `RestatedTrainer`, which restates the same call sequences so that the
GPU box (where the reference does not exist) drives the drop-in through the identical steps:

    _train_one_epoch   train.py:178-241     predict        train.py:129-143
    predict_tiled      train.py:145-176     _save          train.py:320-324
"""
import math
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def reference_dir():
    d = os.environ.get("BRATS_REFERENCE_DIR", "/root/reference")
    return d if os.path.exists(os.path.join(d, "train.py")) else None


def load():
    """(Trainer class, Dice-metric class or None, "reference" | "restated")"""
    d = reference_dir()
    if d is None:
        return RestatedTrainer, RestatedDice, "restated"
    for p in (os.path.join(HERE, "ref_stubs"), d):
        if p not in sys.path:
            sys.path.append(p)          # appended: never shadows a real tensorboardX / nibabel / SimpleITK
    import metrics          # noqa: E402  (the reference's)
    import train            # noqa: E402
    return train.Trainer, metrics.Dice, "reference"


class RestatedState(object):
    def __init__(self):
        self.epoch, self.global_step, self.best_val = 0, 0, 0
        self.train_metric, self.val_metric = dict(), dict()
        self.optimizer_state, self.cuda = None, True


class RestatedDice(object):
    """metrics.py:101-133: thresholded per-sample Dice per channel, both-empty counted as 1, averaged over updates."""

    def __init__(self, name="Dice", input_index=0, target_index=0, classes=4):
        self.name, self.i, self.t, self.classes = name, input_index, target_index, classes
        self.reset()

    def reset(self):
        self.acc, self.samples = 0.0, 0

    def update(self, ground, predict):
        p = (predict[self.i].detach() > 0.5).float().flatten(2)
        g = (ground[self.t].detach() > 0.5).float().flatten(2)
        num, den = (p * g).sum(2), (p + g).sum(2)
        r = torch.where(den > 0, 2 * num / den.clamp_min(1), torch.ones_like(den))[:, :self.classes - 1]
        self.acc = self.acc + r.mean(0).cpu().numpy()
        self.samples += 1

    def get(self):
        return self.acc / max(1, self.samples)


class RestatedTrainer(object):
    def __init__(self, name, models_root, model=None, rewrite=False, connect_tb=True):
        self.model, self.name = model, name
        self.model_path = os.path.join(models_root, name)
        os.makedirs(os.path.join(self.model_path, "logs"), exist_ok=True)
        self.state = RestatedState()
        self.logged = []

    def cuda(self):                                             # train.py:54-57
        self.model.cuda()
        self.state.cuda = True

    def predict(self, batch):                                   # train.py:129-143
        self.model.eval()
        if self.state.cuda:
            self.model.cuda()
        with torch.no_grad():
            assert isinstance(batch[0], list)
            data = [d.cuda() for d in batch[0]] if self.state.cuda else batch[0]
            return self.model(data)

    def predict_tiled(self, batch, output_shape):               # train.py:145-176 with loader_helper.py:34-97
        from oracle import resunet_oracle as O                  # the tiling arithmetic only (checker-side restatement)
        x = batch[0][0]
        out = torch.zeros(output_shape)
        tile, center, border = (192, 192, 192), (48, 48, 48), (72, 72, 72)
        grid = [int(math.ceil(j / i)) for i, j in zip(center, x.shape[2:])]
        for i in range(grid[0]):
            for j in range(grid[1]):
                for k in range(grid[2]):
                    imin, imax = O.tile_get_indices((i, j, k), center, border)
                    t = O.tile_copy(x, tile, imin, imax)
                    if self.state.cuda:
                        t = t.cuda()
                    o = self.model([t])[0].detach().cpu()
                    O.tile_copy_back(out, o, center, imin, imax, border)
        return [out]

    def _train_one_epoch(self, criterion, optimizer, training_data_loader, train_metrics, train_metrics_results, epoch,
                         global_step, scheduler):               # train.py:178-241
        for m in train_metrics:
            m.reset()
        if self.state.cuda:
            self.model.cuda()
        self.model.train()
        optimizer.zero_grad()
        for idx, batch in enumerate(training_data_loader):
            assert isinstance(batch[0], list) and isinstance(batch[1], list)
            data, target = list(batch[0]), list(batch[1])
            if self.state.cuda:
                data, target = [d.cuda() for d in data], [t.cuda() for t in target]
            output = self.model(data)
            loss_val = [c(output, target) for c in criterion]
            loss = sum(loss_val) / len(loss_val)
            loss.backward()
            optimizer.step()
            optimizer.zero_grad()
            if scheduler is not None:
                scheduler.step()
            for m in train_metrics:
                m.update(output, target)
            self.logged.append([l.item() for l in loss_val])    # tb_writer.add_scalar(..., l.item(), ...): host sync per step
            global_step += 1
        for m in train_metrics:
            train_metrics_results[m.name].append(m.get())
        self.state.optimizer_state = optimizer.state_dict()
        return global_step

    def _save(self, suffix):                                    # train.py:320-324: the WHOLE module object is pickled
        torch.save({"state": self.state, "model": self.model}, os.path.join(self.model_path, self.name + suffix + ".pth"))
