"""Stub of tensorboardX for driving the reference's train.py (train.py:7) in tests: SURVEY.md 8c."""


class SummaryWriter(object):
    def __init__(self, logdir=None, **kw):
        self.logdir, self.scalars = logdir, []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))

    def close(self):
        pass
