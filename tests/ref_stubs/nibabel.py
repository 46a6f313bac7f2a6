"""Stub of nibabel for importing the reference's loader_helper.py (loader_helper.py:1) in tests; only the tiling
helpers of that module are used."""


def load(filename):
    raise RuntimeError("nibabel stub: NIfTI reading is outside the hot path")
