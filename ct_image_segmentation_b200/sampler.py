"""GPU-side 3-D patch sampler: random crops of a resident CT volume with fused HU windowing.

Replaces the reference's whole-volume ``Resize3D`` loading (``capstone/volumetric/transforms.py:9-23``,
``capstone/volumetric/datasets.py:24-48``) for training at the named patch sizes (SURVEY.md section
8 f-2).  The volume (int16 HU) and its label map (uint8, 0..9) stay in HBM; one kernel launch
gathers ``batch`` patches, applies ``apply_window`` + normalisation
(``capstone/transforms/transforms_2d.py:97-107``, ``predefined.py:5-29``) and crops the labels.
Patch origins come from a per-rank seeded generator (``seed + rank``), optionally biased towards
foreground voxels; voxels outside the volume are padding (air, label 0).
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, ops
from .transforms import _stacked_window_stats, window_bounds

SEED = 12342  # reference capstone/volumetric/base_trainer.py:18


class PatchSampler:
    def __init__(self, hu: torch.Tensor, labels: Optional[torch.Tensor], patch: Sequence[int],
                 window: str = "soft_tissue", dtype: torch.dtype = torch.float32, rank: int = 0,
                 seed: int = SEED, foreground_prob: float = 0.5, pad_hu: int = -1024):
        if hu.dtype != torch.int16 or hu.dim() != 3 or not hu.is_cuda:
            raise ValueError("hu must be a CUDA int16 (D, H, W) volume")
        if labels is not None and (labels.dtype != torch.uint8 or labels.shape != hu.shape):
            raise ValueError("labels must be uint8 with the volume's shape")
        self.hu, self.labels = hu.contiguous(), None if labels is None else labels.contiguous()
        self.patch = tuple(int(p) for p in patch)
        self.dtype, self.pad_hu = dtype, int(pad_hu)
        self.lo, self.hi = window_bounds(window)
        names = ["brain", "soft_tissue", "bone"]
        self.mean = _stacked_window_stats["mean"][names.index(window)]
        self.std = _stacked_window_stats["std"][names.index(window)]
        self.rng = np.random.default_rng(seed + rank)
        self.foreground_prob = foreground_prob if labels is not None else 0.0
        self._fg = None
        if self.foreground_prob > 0:
            fg = torch.nonzero(self.labels > 0)
            self._fg = fg.cpu().numpy() if fg.numel() else None

    def origins(self, batch: int) -> np.ndarray:
        """(batch, 3) int32 patch origins: uniform, or centred on a random foreground voxel."""
        dims = np.asarray(self.hu.shape)
        p = np.asarray(self.patch)
        out = np.empty((batch, 3), dtype=np.int32)
        for b in range(batch):
            if self._fg is not None and self.rng.random() < self.foreground_prob:
                centre = self._fg[self.rng.integers(len(self._fg))]
                o = centre - p // 2
            else:
                o = np.array([self.rng.integers(0, max(d - q, 0) + 1) for d, q in zip(dims, p)])
            out[b] = np.clip(o, np.minimum(dims - p, 0), np.maximum(dims - p, 0))
        return out

    def sample(self, batch: int) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """(images (B, 1, *patch) ``dtype``, labels (B, *patch) uint8 or None)."""
        lib = _lib.load()
        org = torch.from_numpy(self.origins(batch)).to(self.hu.device)
        pd, ph, pw = self.patch
        img = torch.empty(batch, 1, pd, ph, pw, dtype=self.dtype, device=self.hu.device)
        lab = None if self.labels is None else torch.empty(batch, pd, ph, pw, dtype=torch.uint8, device=self.hu.device)
        d, h, w = self.hu.shape
        _lib.check(lib.b200seg_crop_window_norm(
            ops.dtype_code(self.dtype), self.hu.data_ptr(), None if self.labels is None else self.labels.data_ptr(),
            org.data_ptr(), batch, img.data_ptr(), None if lab is None else lab.data_ptr(), d, h, w, pd, ph, pw,
            float(self.lo), float(self.hi), float(self.mean), float(self.std), self.pad_hu,
            torch.cuda.current_stream().cuda_stream), "b200seg_crop_window_norm")
        self.last_origins = org
        return img, lab
