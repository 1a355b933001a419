"""On-disk formats either side of the hot path.

* Pre-processed patients: the ``.npz`` files the reference writes with ``np.savez(image=..., masks=...,
  mask_indicator=...)`` (``capstone/data/process_miccai.py:96-131``): ``image`` (1, D, H, W) HU volume
  (or (1, H, W) for a 2-D slice, ``:60-93``), ``masks`` (9, D, H, W) / (9, H, W) uint8 -- one binary mask per
  structure in ``STRUCTURES`` order -- and ``mask_indicator`` (9,) with 0 for structures that were not
  annotated.  ``PatientVolume.sampler`` puts the volume in HBM (int16 HU + squashed uint8 label map, fused
  ``b200seg_squash_masks``) behind the GPU patch sampler.
* Checkpoints: PyTorch-Lightning ``.ckpt`` files of the reference's ``BaseUNet3D`` / ``BaseUNet2D``
  (``capstone/paths.py:46-49``, loaded at ``capstone/interpretability.py:28-31``): a dict with
  ``state_dict`` whose U-Net keys carry the ``unet.`` prefix and ``hyper_parameters``.
"""
from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional, Sequence, Union

import numpy as np
import torch

from .losses import STRUCTURES


@dataclass
class PatientVolume:
    image: np.ndarray            # (D, H, W) HU, int16
    masks: np.ndarray            # (9, D, H, W) uint8
    mask_indicator: np.ndarray   # (9,) float32, 0 = structure not annotated
    patient_id: str = ""

    def labels(self) -> np.ndarray:
        """(D, H, W) uint8 label map, ``max_c mask_c * (c + 1)`` (reference ``capstone/volumetric/utils.py:4-7``)."""
        ids = np.arange(1, self.masks.shape[0] + 1, dtype=np.uint8).reshape(-1, 1, 1, 1)
        return (self.masks.astype(np.uint8) * ids).max(axis=0)

    def sampler(self, patch: Sequence[int], device="cuda", **kwargs):
        """GPU patch sampler over this volume (volume + label map resident in HBM)."""
        from . import metrics
        from .sampler import PatchSampler
        hu = torch.from_numpy(self.image).to(device)
        masks = torch.from_numpy(self.masks).to(device).unsqueeze(0)
        labels = metrics.squash_masks(masks)[0].to(torch.uint8)
        return PatchSampler(hu, labels, patch, **kwargs)


def load_patient_npz(path: Union[str, Path]) -> PatientVolume:
    """Read one pre-processed patient (3-D) or slice (2-D, returned with D = 1)."""
    path = Path(path)
    with np.load(path) as z:
        missing = {"image", "masks", "mask_indicator"} - set(z.files)
        if missing:
            raise KeyError(f"{path}: not a reference .npz (missing {sorted(missing)})")
        image, masks, ind = z["image"], z["masks"], z["mask_indicator"]
    if masks.ndim == 3:
        # 2-D slice files (``_patient_to_2d``, ``capstone/data/process_miccai.py:60-93``): ``image`` is
        # ``vol[:, index]`` = (1, H, W) -- the leading axis is the CHANNEL, not a depth -- and ``masks`` (9, H, W)
        if image.ndim == 3 and image.shape[0] == 1:
            image = image[0]
        if image.ndim != 2:
            raise ValueError(f"{path}: 2-D masks {masks.shape} need an (H, W) or (1, H, W) image, got {image.shape}")
        image, masks = image[None], masks[:, None]
    elif image.ndim == 4 and image.shape[0] == 1:
        image = image[0]
    if image.ndim != 3 or masks.ndim != 4 or masks.shape[1:] != image.shape:
        raise ValueError(f"{path}: image {image.shape} / masks {masks.shape} are not (D, H, W) / (S, D, H, W)")
    if masks.shape[0] != len(STRUCTURES) or ind.shape != (len(STRUCTURES),):
        raise ValueError(f"{path}: expected {len(STRUCTURES)} structures, got masks {masks.shape[0]}, indicator {ind.shape}")
    hu = np.clip(np.rint(image.astype(np.float64)), -32768, 32767).astype(np.int16)
    return PatientVolume(np.ascontiguousarray(hu), np.ascontiguousarray((masks != 0).astype(np.uint8)),
                         ind.astype(np.float32), path.stem)


def save_patient_npz(path: Union[str, Path], image: np.ndarray, masks: np.ndarray,
                     mask_indicator: Optional[np.ndarray] = None) -> None:
    """Write the reference's layout (``image`` gets the leading singleton channel the 3-D files have)."""
    image = np.asarray(image)
    if image.ndim == 3:
        image = image[None]
    if mask_indicator is None:
        mask_indicator = np.ones(len(STRUCTURES))
    np.savez(str(path), image=image, masks=np.asarray(masks), mask_indicator=np.asarray(mask_indicator))


def _load_checkpoint_file(path: Union[str, Path], allow_pickle: bool = False) -> Dict:
    """``torch.load`` restricted to tensors and plain containers (``weights_only=True``).  Lightning checkpoints
    whose ``hyper_parameters`` hold arbitrary pickled objects need the explicit opt-in ``allow_pickle=True``
    (arbitrary code execution on load: only for files you trust)."""
    try:
        return torch.load(str(path), map_location="cpu", weights_only=True)
    except Exception as e:  # noqa: BLE001
        if not allow_pickle:
            raise RuntimeError(f"{path}: not loadable with weights_only=True ({e}); pass allow_pickle=True "
                               f"if the file is trusted") from e
        return torch.load(str(path), map_location="cpu", weights_only=False)


def unet_state_dict_from_checkpoint(ckpt: Union[str, Path, Dict], allow_pickle: bool = False) -> Dict[str, torch.Tensor]:
    """The U-Net's ``state_dict`` out of a reference Lightning checkpoint (file or already-loaded dict):
    keys ``unet.model...`` -> ``model...``; a plain U-Net state dict passes through."""
    if not isinstance(ckpt, dict):
        ckpt = _load_checkpoint_file(ckpt, allow_pickle)
    sd = ckpt.get("state_dict", ckpt)
    out = {k[len("unet."):]: v for k, v in sd.items() if k.startswith("unet.")}
    if not out:
        out = {k: v for k, v in sd.items() if k.startswith("model.")}
    if not out:
        raise KeyError("no U-Net weights ('unet.model.*' or 'model.*') in the checkpoint")
    return out


def load_lightning_checkpoint(ckpt: Union[str, Path, Dict], net: torch.nn.Module, allow_pickle: bool = False) -> Dict:
    """Load the reference's released weights (``model_large.ckpt`` / ``model_mixup.ckpt``) into a b200seg
    ``UNet`` built with the same hyper-parameters; returns the checkpoint's ``hyper_parameters``."""
    loaded = ckpt if isinstance(ckpt, dict) else _load_checkpoint_file(ckpt, allow_pickle)
    net.load_state_dict(unet_state_dict_from_checkpoint(loaded), strict=True)
    return dict(loaded.get("hyper_parameters", {})) if isinstance(loaded, dict) else {}
