"""Reference ``capstone/volumetric/utils.py:4-7``."""
from ...metrics import squash_masks as _squash_masks_impl


def _squash_masks_3D(masks, n_classes, device=None):
    return _squash_masks_impl(masks, n_classes)
