"""b200seg: B200-native (sm_100a) U-Net + softmax-Dice hot path behind the reference's API.

Drop-in surface (reference ``capstone/models/__init__.py:1-3``)::

    from ct_image_segmentation_b200 import UNet, MultipleLossWrapper, DiceMetricWrapper

The CUDA library (``lib/libb200seg.so``) is loaded lazily on first use; there is no CPU or
PyTorch fallback for the hot path.
"""
from .losses import (CrossEntropyLoss, DiceLoss, DiceLossWrapper, FocalLoss, GeneralizedDiceLoss,
                     MultipleLossWrapper, MultipleLossWrapper3D, N_CLASSES, STRUCTURES, apply_missing_mask)
from .metrics import (DiceMetricWrapper, DiceMetricWrapper3D, dice_from_counts, squash_masks,
                      squash_predictions)
from .unet import UNet
from .engine import GraphedTrainStep
from .optim import FlatAdam

__all__ = [
    "UNet", "GraphedTrainStep", "FlatAdam", "DiceLoss", "GeneralizedDiceLoss", "FocalLoss", "CrossEntropyLoss", "DiceLossWrapper", "MultipleLossWrapper",
    "MultipleLossWrapper3D", "DiceMetricWrapper", "DiceMetricWrapper3D", "apply_missing_mask",
    "squash_masks", "squash_predictions", "dice_from_counts", "STRUCTURES", "N_CLASSES",
]
