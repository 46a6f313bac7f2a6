"""B200-native (sm_100a) implementation of the lachinov/brats2019 ResUNet hot path.

    from brats2019_b200 import UNet, Dice_loss_joint      # drop-ins for model.UNet / loss.Dice_loss_joint

Importing the package does not load the CUDA library; the first kernel call does, and raises
if `libbrats_b200.so` has not been built (there is no fallback path).
"""
from .loss import BCE_Loss, Dice_loss_joint  # noqa: F401
from .model import Residual, Trilinear, UNet, conv  # noqa: F401

DEFAULT_CFG = dict(depth=4, encoder_layers=[1, 2, 2, 4], decoder_layers=[1, 1, 1, 1],
                   number_of_channels=[16, 32, 64, 128], number_of_outputs=3)   # main.py:56-59

__all__ = ["UNet", "Residual", "conv", "Trilinear", "Dice_loss_joint", "BCE_Loss", "DEFAULT_CFG"]
