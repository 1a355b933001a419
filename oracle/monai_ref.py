"""CPU oracle: restatement of the reference hot path on ``torch.nn`` primitives.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

What is restated
----------------
The reference's hot-path arithmetic is not in its repository: ``UNet`` *is*
``monai.networks.nets.UNet`` (reference ``capstone/models/__init__.py:3``) and
the Dice loss is ``monai.losses.DiceLoss`` (``capstone/models/losses.py:3,80-85``),
MONAI 0.3 per the reference ``README.md:39``.  MONAI is not installed here and
is not vendored in ``/root/reference``, so this file restates MONAI-0.3's
published structure using only ``torch.nn`` modules, so
that torch itself is the arithmetic oracle.

Pinning status
--------------
* ``UNet`` structure: pinned by the only architectural known-answers the
  reference publishes -- parameter counts 26 M / 13.5 M (``reports/Report.pdf``
  Table 1; exact values in SURVEY.md A.5) and the module path
  ``unet.model[2][1].conv.unit0.conv`` (``capstone/interpretability.py:88``).
  Numerics beyond that are torch's own Conv/InstanceNorm/PReLU:
  **parity unpinned** by any reference golden vector (the reference has no tests).
* ``DiceLoss``: pinned against the reference's in-tree
  ``GeneralizedDiceLoss(w_type="uniform")`` (``capstone/models/temp.py:17-170``),
  which is the same formula; fixtures in ``tests/golden/`` are produced by
  importing that file (``tests/golden/make_golden.py``).
* Dice metric / reductions / mask squashing / missing-annotation masking /
  HU windowing: pinned against the reference's own functions the same way.
* ``GeneralizedDiceLoss`` ("square" / "simple" weights incl. the inf -> max rule) and ``BoundaryLoss``: pinned
  against the reference's own ``capstone/models/temp.py`` / ``capstone/models/losses.py`` classes the same way.
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

# reference capstone/utils/miccai.py:14-24 -- order is load-bearing (class id = index + 1)
STRUCTURES = [
    "BrainStem",
    "Chiasm",
    "Mandible",
    "OpticNerve_L",
    "OpticNerve_R",
    "Parotid_L",
    "Parotid_R",
    "Submandibular_L",
    "Submandibular_R",
]
N_CLASSES = len(STRUCTURES) + 1

_CONV = {1: nn.Conv1d, 2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {1: nn.ConvTranspose1d, 2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_INORM = {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}


# ----------------------------------------------------------------------------
# MONAI-0.3 network blocks
# ----------------------------------------------------------------------------
class Convolution(nn.Sequential):
    """conv -> InstanceNorm(affine=False) -> PReLU; ``conv_only`` keeps just the conv.

    MONAI 0.3 ``monai.networks.blocks.Convolution`` with the arguments the
    reference's UNet uses: same padding ``(k-1)//2``, bias,
    transposed variant has ``output_padding = stride - 1``.
    """

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3,
                 conv_only=False, is_transposed=False):
        super().__init__()
        pad = (kernel_size - 1) // 2
        if is_transposed:
            conv = _CONVT[dimensions](in_channels, out_channels, kernel_size, stride=strides,
                                      padding=pad, output_padding=strides - 1, bias=True)
        else:
            conv = _CONV[dimensions](in_channels, out_channels, kernel_size, stride=strides,
                                     padding=pad, bias=True)
        self.add_module("conv", conv)
        if not conv_only:
            self.add_module("norm", _INORM[dimensions](out_channels))
            self.add_module("act", nn.PReLU())


class ResidualUnit(nn.Module):
    """``conv(x) + residual(x)``."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3,
                 subunits=2, last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        subunits = max(1, subunits)
        sc, ss = in_channels, strides
        for su in range(subunits):
            only = last_conv_only and su == subunits - 1
            self.conv.add_module(
                f"unit{su:d}",
                Convolution(dimensions, sc, out_channels, strides=ss, kernel_size=kernel_size,
                            conv_only=only),
            )
            sc, ss = out_channels, 1
        if strides != 1 or in_channels != out_channels:
            rk, rp = kernel_size, (kernel_size - 1) // 2
            if strides == 1:  # only the channel count changes: 1x1 conv, no padding
                rk, rp = 1, 0
            self.residual = _CONV[dimensions](in_channels, out_channels, rk, strides, rp, bias=True)

    def forward(self, x):
        return self.conv(x) + self.residual(x)


class SkipConnection(nn.Module):
    """``cat([x, submodule(x)], dim=1)`` -- x first."""

    def __init__(self, submodule):
        super().__init__()
        self.submodule = submodule

    def forward(self, x):
        return torch.cat([x, self.submodule(x)], dim=1)


class UNet(nn.Module):
    """Restated ``monai.networks.nets.UNet`` (0.3); SURVEY.md A.1.

    Constructed by the reference at ``capstone/volumetric/base_trainer.py:65-72``
    and ``capstone/training/base_trainer.py:72-79``.
    """

    def __init__(self, dimensions, in_channels, out_channels, channels: Sequence[int],
                 strides: Sequence[int], kernel_size=3, up_kernel_size=3, num_res_units=0):
        super().__init__()
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.channels = list(channels)
        self.strides = list(strides)
        self.kernel_size = kernel_size
        self.up_kernel_size = up_kernel_size
        self.num_res_units = num_res_units

        def block(inc, outc, chans, strds, is_top):
            c, s = chans[0], strds[0]
            if len(chans) > 2:
                sub = block(c, c, chans[1:], strds[1:], False)
                upc = c * 2
            else:
                sub = self._down(c, chans[1], 1)
                upc = c + chans[1]
            return nn.Sequential(self._down(inc, c, s), SkipConnection(sub),
                                 self._up(upc, outc, s, is_top))

        self.model = block(in_channels, out_channels, self.channels, self.strides, True)

    def _down(self, inc, outc, s):
        if self.num_res_units > 0:
            return ResidualUnit(self.dimensions, inc, outc, strides=s,
                                kernel_size=self.kernel_size, subunits=self.num_res_units)
        return Convolution(self.dimensions, inc, outc, strides=s, kernel_size=self.kernel_size)

    def _up(self, inc, outc, s, is_top):
        conv = Convolution(self.dimensions, inc, outc, strides=s, kernel_size=self.up_kernel_size,
                           conv_only=is_top and self.num_res_units == 0, is_transposed=True)
        if self.num_res_units > 0:
            ru = ResidualUnit(self.dimensions, outc, outc, strides=1, kernel_size=self.kernel_size,
                              subunits=1, last_conv_only=is_top)
            return nn.Sequential(conv, ru)
        return conv

    def forward(self, x):
        return self.model(x)


def count_parameters(m: nn.Module) -> int:
    return sum(p.numel() for p in m.parameters())


# ----------------------------------------------------------------------------
# Losses (reference capstone/models/losses.py)
# ----------------------------------------------------------------------------
def one_hot(labels: torch.Tensor, num_classes: int) -> torch.Tensor:
    """``monai.networks.one_hot``: (B,1,*S) integer labels -> (B,C,*S) float32."""
    assert labels.shape[1] == 1, "labels must have a singleton channel dim"
    shape = list(labels.shape)
    shape[1] = num_classes
    out = torch.zeros(shape, dtype=torch.float32, device=labels.device)
    return out.scatter_(1, labels.long(), 1.0)


class DiceLoss(nn.Module):
    """``monai.losses.DiceLoss`` (0.3) as the reference configures it
    (``capstone/models/losses.py:78-85``, ``capstone/volumetric/losses.py:70-77``)."""

    def __init__(self, include_background=True, to_onehot_y=False, softmax=False,
                 reduction="mean", smooth=1e-5):
        super().__init__()
        self.include_background = include_background
        self.to_onehot_y = to_onehot_y
        self.softmax = softmax
        self.reduction = reduction
        self.smooth = smooth

    def forward(self, input, target):
        c = input.shape[1]
        p = torch.softmax(input, 1) if (self.softmax and c > 1) else input
        t = one_hot(target, c) if (self.to_onehot_y and c > 1) else target
        if not self.include_background and c > 1:
            p, t = p[:, 1:], t[:, 1:]
        assert p.shape == t.shape
        axes = list(range(2, p.ndim))
        inter = (t * p).sum(axes)
        denom = t.sum(axes) + p.sum(axes)
        f = 1.0 - (2.0 * inter + self.smooth) / (denom + self.smooth)
        if self.reduction == "mean":
            return f.mean()
        if self.reduction == "sum":
            return f.sum()
        if self.reduction == "none":
            return f
        raise ValueError(self.reduction)


class GeneralizedDiceLoss(nn.Module):
    """The reference's in-tree ``GeneralizedDiceLoss`` (``capstone/models/temp.py:17-170``; softmax, one-hot target,
    ``batch=False``): per (n, c) ``1 - (2 I w + s) / ((G + P) w + s)`` with ``w = 1/G^2`` ("square"), ``1/G``
    ("simple") or 1 ("uniform"); infinite weights (class absent from the sample) become the largest finite weight
    of that sample (``:146-152``).  Pinned by fixtures made from that file (``tests/golden/make_golden.py``)."""

    def __init__(self, include_background=True, to_onehot_y=False, softmax=False, w_type="square",
                 reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5):
        super().__init__()
        self.include_background, self.to_onehot_y, self.softmax = include_background, to_onehot_y, softmax
        self.w_type, self.reduction, self.smooth_nr, self.smooth_dr = w_type, reduction, smooth_nr, smooth_dr

    def forward(self, input, target):
        c = input.shape[1]
        p = torch.softmax(input, 1) if (self.softmax and c > 1) else input
        t = one_hot(target, c) if (self.to_onehot_y and c > 1) else target
        if not self.include_background and c > 1:
            p, t = p[:, 1:], t[:, 1:]
        axes = list(range(2, p.ndim))
        inter, ground, pred = (t * p).sum(axes), t.sum(axes), p.sum(axes)
        g = ground.float()
        w = {"simple": torch.reciprocal(g), "square": torch.reciprocal(g * g)}.get(self.w_type, torch.ones_like(g))
        for row in w:
            infs = torch.isinf(row)
            row[infs] = 0.0
            row[infs] = torch.max(row)
        f = 1.0 - (2.0 * (inter * w) + self.smooth_nr) / ((ground + pred) * w + self.smooth_dr)
        if self.reduction == "mean":
            return f.mean()
        if self.reduction == "sum":
            return f.sum()
        return f


class BoundaryLoss(nn.Module):
    """``BoundaryLossWrapper`` of the reference (``capstone/models/losses.py:127-157``): ``softmax(input)[:, 1:] *
    dist_maps``, mean over everything or over the spatial axes.  Pinned by fixtures made from that class."""

    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction

    def forward(self, input, dist_maps):
        loss = torch.softmax(input, dim=1)[:, 1:] * dist_maps.type_as(input)
        if self.reduction == "none":
            return loss.mean(dim=tuple(range(2, loss.ndim)))
        return loss.mean()


class FocalLoss(nn.Module):
    """``monai.losses.FocalLoss`` (0.3), gamma=2, one-hot target path
    (reference ``capstone/models/losses.py:105-124``).  **parity unpinned**."""

    def __init__(self, gamma=2.0, reduction="mean"):
        super().__init__()
        self.gamma = gamma
        self.reduction = reduction

    def forward(self, input, target):
        b, c = input.shape[:2]
        i = input.reshape(b, c, -1)
        t = target.reshape(b, target.shape[1], -1)
        logpt = F.log_softmax(i, dim=1)
        if t.shape[1] == 1:
            logpt = logpt.gather(1, t.long()).squeeze(1)
        pt = logpt.exp()
        w = (1.0 - pt) ** self.gamma
        if t.shape[1] == 1:
            loss = (-w * logpt).mean(dim=1)
        else:
            loss = (-w * t * logpt).mean(dim=-1)
        if self.reduction == "mean":
            return loss.mean()
        if self.reduction == "sum":
            return loss.sum()
        return loss


def apply_missing_mask(name, loss, mask_indicator):
    """AnatomyNet missing-annotation weighting; reference ``capstone/models/losses.py:206-221``."""
    if name == "Focal":
        bg = (mask_indicator.sum(dim=1, keepdim=True) == (N_CLASSES - 1)).float()
        mask_indicator = torch.cat([bg, mask_indicator], dim=1)
    w = 1.0 / mask_indicator.sum(dim=0)
    if torch.isinf(w).any():
        w = torch.ones_like(w)
    w = w / w.sum()
    return (loss * w[None, :] * mask_indicator).sum(dim=1).mean()


class MultipleLossWrapper(nn.Module):
    """Intended 3D behaviour of reference ``capstone/models/losses.py:170-203`` /
    ``capstone/volumetric/losses.py:128-130`` (no ndim asserts)."""

    def __init__(self, losses, exclude_missing=False):
        super().__init__()
        self.names = list(losses)
        self.exclude_missing = exclude_missing
        self.reduction = "none" if exclude_missing else "mean"

    def forward(self, input, target, mask_indicator=None, dist_maps=None):
        out = {}
        if mask_indicator is not None:
            mask_indicator = mask_indicator.type_as(input)
        for name in self.names:
            if name == "Boundary":
                assert dist_maps is not None, "Distance maps are required for using boundary loss"
                v = BoundaryLoss(reduction=self.reduction)(input, dist_maps)
            elif name == "GeneralizedDice":
                v = GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True,
                                        reduction=self.reduction)(input, target.unsqueeze(1))
            elif name == "Dice":
                v = DiceLoss(include_background=False, to_onehot_y=True, softmax=True,
                             reduction=self.reduction)(input, target.unsqueeze(1))
            elif name == "Focal":
                v = FocalLoss(reduction=self.reduction)(input, one_hot(target.unsqueeze(1), N_CLASSES))
            elif name == "CrossEntropy":
                v = F.cross_entropy(input, target)
            else:
                raise KeyError(name)
            if self.exclude_missing and name not in ("CrossEntropy", "WeightedCrossEntropy"):
                v = apply_missing_mask(name, v, mask_indicator)
            out[name] = v
        return out


# ----------------------------------------------------------------------------
# Label maps and the Dice metric
# ----------------------------------------------------------------------------
def squash_masks(masks: torch.Tensor, n_classes: int = N_CLASSES) -> torch.Tensor:
    """(B,9,*S) binary masks -> (B,*S) label map, ``max_c mask_c*(c+1)``.
    Reference ``capstone/volumetric/utils.py:4-7``, ``capstone/training/utils.py:13-16``."""
    ids = torch.arange(1, n_classes, device=masks.device)
    shape = [1, -1] + [1] * (masks.ndim - 2)
    return (masks * ids.view(shape)).max(dim=1).values


def squash_predictions(preds: torch.Tensor) -> torch.Tensor:
    """softmax(dim=1) then argmax(dim=1); reference ``capstone/training/utils.py:19-20``."""
    return torch.softmax(preds, dim=1).argmax(dim=1)


def dice_counts(pred: torch.Tensor, target: torch.Tensor, n_classes: int = N_CLASSES):
    """Integer form of the metric's sums: per (b, c) TP, |pred==c|, |target==c|."""
    b = pred.shape[0]
    p = pred.reshape(b, -1).long()
    t = target.reshape(b, -1).long()
    tp = torch.zeros(b, n_classes, dtype=torch.int64)
    np_ = torch.zeros(b, n_classes, dtype=torch.int64)
    nt = torch.zeros(b, n_classes, dtype=torch.int64)
    for i in range(b):
        np_[i] = torch.bincount(p[i], minlength=n_classes)[:n_classes]
        nt[i] = torch.bincount(t[i], minlength=n_classes)[:n_classes]
        tp[i] = torch.bincount(p[i][p[i] == t[i]], minlength=n_classes)[:n_classes]
    return tp, np_, nt


def dice_metric(pred: torch.Tensor, target: torch.Tensor, n_classes: int = N_CLASSES):
    """``DiceMetricWrapper(3D).__call__``: reference ``capstone/models/metrics.py:15-21``
    -> ``compute_meandice(include_background=False)`` (``capstone/models/temp.py:173-214``)
    -> ``do_metric_reduction("mean_batch")`` (``temp.py:233-292``).
    Returns (mean over the 9 classes, per-class[9])."""
    tp, np_, nt = dice_counts(pred, target, n_classes)
    tp, np_, nt = tp[:, 1:].float(), np_[:, 1:].float(), nt[:, 1:].float()
    score = torch.where(nt > 0, 2.0 * tp / (nt + np_), torch.full_like(tp, float("nan")))
    valid = ~torch.isnan(score)
    cnt = valid.float().sum(dim=0)
    tot = torch.where(valid, score, torch.zeros_like(score)).sum(dim=0)
    per_class = torch.where(cnt > 0, tot / cnt, torch.zeros_like(tot))
    return per_class.mean(), per_class


# ----------------------------------------------------------------------------
# HU windowing + normalisation (reference capstone/transforms)
# ----------------------------------------------------------------------------
WINDOWING_CONFIG = {"brain": (80, 40), "soft_tissue": (350, 20), "bone": (2800, 600)}
WINDOW_MEAN = (0.107, 0.135, 0.085)
WINDOW_STD = (0.271, 0.267, 0.152)


def apply_window(image: np.ndarray, width: int, level: int, shift: bool = True) -> np.ndarray:
    """Reference ``capstone/transforms/transforms_2d.py:97-107`` (float64 arithmetic for
    integer inputs, as numpy promotes)."""
    lo = level - (width // 2)
    hi = level + (width // 2)
    out = np.clip(image, lo, hi)
    if shift:
        out = (out - lo) / (hi - lo + 1e-8)
    return out


def window_normalize(image: np.ndarray, windows=("soft_tissue",), mean=None, std=None) -> np.ndarray:
    """``WindowedChannels``/``SoftTissueWindowing`` followed by albumentations
    ``Normalize(mean, std, max_pixel_value=1.0)`` (reference
    ``capstone/transforms/predefined.py:5-29``): channels-last float32 output."""
    if mean is None:
        mean = WINDOW_MEAN if len(windows) == 3 else (WINDOW_MEAN[1],)
    if std is None:
        std = WINDOW_STD if len(windows) == 3 else (WINDOW_STD[1],)
    chans = [apply_window(image, *WINDOWING_CONFIG[w]) for w in windows]
    x = np.stack(chans, axis=-1).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)
    s = np.asarray(std, dtype=np.float32)
    # albumentations.normalize: img = img.astype(float32); img -= mean*max; img *= 1/(std*max)
    x = x - m
    x = x * np.reciprocal(s, dtype=np.float32)
    return x


# ----------------------------------------------------------------------------
# Sliding-window inference (MONAI semantics) -- spec the build owns
# ----------------------------------------------------------------------------
def _scan_starts(size, roi, overlap):
    if roi >= size:
        return [0]
    interval = int(roi * (1 - overlap))
    interval = max(interval, 1)
    n = 1
    while (n - 1) * interval + roi < size:
        n += 1
    return [min(k * interval, size - roi) for k in range(n)]


def importance_map(roi_size, mode="constant", sigma_scale=0.125):
    """MONAI ``compute_importance_map``: ones, or a Gaussian centred on the window (sigma = sigma_scale * roi per
    axis), normalised to max 1 with its zeros lifted to the smallest positive value.
    **parity unpinned** (MONAI semantics from memory)."""
    if mode == "constant":
        return torch.ones(tuple(roi_size), dtype=torch.float32)
    if mode != "gaussian":
        raise ValueError(mode)
    imp = torch.ones((), dtype=torch.float64)
    for k, r in enumerate(roi_size):
        x = torch.arange(r, dtype=torch.float64) - (r // 2)  # a unit impulse at r // 2, Gaussian-filtered
        g = torch.exp(-0.5 * (x / (sigma_scale * r)) ** 2)
        shape = [1] * len(roi_size)
        shape[k] = r
        imp = imp * g.reshape(shape)
    imp = imp / imp.max()
    imp = imp.float()
    pos = imp[imp > 0]
    return torch.clamp(imp, min=float(pos.min()))


def sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25, mode="constant"):
    """Sliding window with constant or Gaussian importance; batch 1; **parity unpinned** (absent from the
    reference, F6)."""
    assert inputs.shape[0] == 1
    imp = importance_map(roi_size, mode)
    dims = inputs.shape[2:]
    pad = []
    for sz, r in zip(reversed(dims), reversed(roi_size)):
        diff = max(r - sz, 0)
        pad += [diff // 2, diff - diff // 2]
    x = F.pad(inputs, pad) if any(pad) else inputs
    pdims = x.shape[2:]
    starts = [_scan_starts(s, r, overlap) for s, r in zip(pdims, roi_size)]
    wins = [(a, b, c) for a in starts[0] for b in starts[1] for c in starts[2]]
    out = None
    cnt = torch.zeros((1, 1) + tuple(pdims), dtype=torch.float32)
    for i in range(0, len(wins), sw_batch_size):
        chunk = wins[i:i + sw_batch_size]
        batch = torch.cat([x[:, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]]
                           for a, b, c in chunk], 0)
        pred = predictor(batch).float()
        if out is None:
            out = torch.zeros((1, pred.shape[1]) + tuple(pdims), dtype=torch.float32)
        for j, (a, b, c) in enumerate(chunk):
            out[:, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]] += pred[j:j + 1] * imp
            cnt[:, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]] += imp
    out = out / cnt
    sl = [slice(None), slice(None)]
    for k, (sz, r) in enumerate(zip(dims, roi_size)):
        diff = max(r - sz, 0)
        sl.append(slice(diff // 2, diff // 2 + sz))
    return out[tuple(sl)]


# ----------------------------------------------------------------------------
# Reference training step (fwd + bwd + Dice), used as the CPU baseline
# ----------------------------------------------------------------------------
def train_step(unet: nn.Module, images: torch.Tensor, labels: torch.Tensor):
    """forward -> softmax Dice (reduction mean) -> backward, as
    ``BaseUNet3D._shared_step`` + ``loss.backward()`` do (reference
    ``capstone/volumetric/base_trainer.py:87-111``)."""
    for p in unet.parameters():
        p.grad = None
    logits = unet(images)
    loss = DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(
        logits, labels.unsqueeze(1))
    loss.backward()
    return logits.detach(), loss.detach()
