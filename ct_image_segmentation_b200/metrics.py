"""Label maps and the Dice metric on fused integer-counting kernels.

Reference: ``_squash_predictions`` (``capstone/training/utils.py:19-20``), ``_squash_masks[_3D]``
(``capstone/training/utils.py:13-16``, ``capstone/volumetric/utils.py:4-7``) and
``DiceMetricWrapper[3D]`` (``capstone/models/metrics.py:8-31``, ``capstone/volumetric/metrics.py:5-20``)
-> ``compute_meandice`` / ``do_metric_reduction("mean_batch")`` (``capstone/models/temp.py:173-292``).
The reference materialises two (B, 10, *S) fp32 one-hot tensors; here one pass reads the label
maps (or the logits) and produces 3*B*C exact integer counts.
"""
from __future__ import annotations

import torch

from . import ops
from .losses import N_CLASSES, _as_cl


def squash_predictions(preds: torch.Tensor) -> torch.Tensor:
    """softmax(dim=1).argmax(dim=1), first maximum wins: (B, C, *S) -> (B, *S) int64."""
    cl = _as_cl(preds)
    pred, _ = ops.argmax_dice_counts(cl, None, want_pred=True)
    pred = pred.squeeze(1) if preds.dim() == 4 else pred
    return pred.long()


def squash_masks(masks: torch.Tensor, n_classes: int = N_CLASSES, device=None) -> torch.Tensor:
    """(B, 9, *S) binary masks -> (B, *S) label map (uint8 values 0..9)."""
    if masks.shape[1] != n_classes - 1:
        raise ValueError(f"expected {n_classes - 1} structure masks, got {masks.shape[1]}")
    return ops.squash_masks(masks)


def dice_from_counts(counts: torch.Tensor):
    """(B, C, 3) integer counts {tp, |pred|, |target|} -> (mean, per-class[C-1]) exactly as
    compute_meandice(include_background=False) + do_metric_reduction("mean_batch") + .mean()."""
    c = counts[:, 1:].to(torch.float32)
    tp, n_pred, n_tgt = c[..., 0], c[..., 1], c[..., 2]
    has = n_tgt > 0
    score = torch.where(has, 2.0 * tp / (n_tgt + n_pred).clamp_min(1.0), torch.zeros_like(tp))
    cnt = has.float().sum(dim=0)
    per_class = torch.where(cnt > 0, score.sum(dim=0) / cnt.clamp_min(1.0), torch.zeros_like(cnt))
    return per_class.mean(), per_class


class DiceMetricWrapper(object):
    """``DiceMetricWrapper()(pred_labels, target_labels) -> (dice_mean, dice_per_class[9])``."""

    def __init__(self):
        self.n_classes = N_CLASSES

    def __call__(self, input: torch.Tensor, target: torch.Tensor):
        counts = ops.label_dice_counts(input, target, self.n_classes)
        return dice_from_counts(counts)

    def from_logits(self, logits: torch.Tensor, target: torch.Tensor):
        """Fused ``_squash_predictions`` + metric: one read of the logits."""
        _, counts = ops.argmax_dice_counts(_as_cl(logits), target, want_pred=False)
        return dice_from_counts(counts)


DiceMetricWrapper3D = DiceMetricWrapper
