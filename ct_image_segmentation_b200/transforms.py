"""HU windowing + normalisation as one fused elementwise kernel.

Reference: ``apply_window`` / ``WINDOWING_CONFIG`` (``capstone/transforms/transforms_2d.py:6,97-107``),
``WindowedChannels`` (``:9-39``) and the albumentations ``Normalize`` constants of
``capstone/transforms/predefined.py:5-29``.
"""
from __future__ import annotations

from typing import Sequence

import torch

from . import ops

WINDOWING_CONFIG = {"brain": (80, 40), "soft_tissue": (350, 20), "bone": (2800, 600)}
_stacked_window_stats = {"mean": (0.107, 0.135, 0.085), "std": (0.271, 0.267, 0.152)}


def window_bounds(name: str):
    width, level = WINDOWING_CONFIG[name]
    return level - (width // 2), level + (width // 2)


def window_normalize(hu: torch.Tensor, windows: Sequence[str] = ("soft_tissue",), mean=None, std=None,
                     dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """int16 HU tensor (any shape) -> (*shape, len(windows)) channels-last windowed, shifted to
    [0, 1] and normalised with the reference's statistics."""
    names = list(WINDOWING_CONFIG)
    if mean is None:
        mean = [_stacked_window_stats["mean"][names.index(w)] for w in windows]
    if std is None:
        std = [_stacked_window_stats["std"][names.index(w)] for w in windows]
    lo, hi = zip(*[window_bounds(w) for w in windows])
    return ops.hu_window_norm(hu, lo, hi, mean, std, dtype)
