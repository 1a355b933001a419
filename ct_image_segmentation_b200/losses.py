"""Softmax Dice loss and the reference's loss wrappers on the fused sm_100a kernels.

Replaces ``monai.losses.DiceLoss`` as configured by the reference
(``capstone/models/losses.py:71-85``, ``capstone/volumetric/losses.py:63-77``) and keeps the
reference's wrapper API: ``MultipleLossWrapper(losses, exclude_missing)(input, target,
mask_indicator)`` (``capstone/models/losses.py:170-203``) with AnatomyNet missing-annotation
weighting (``:206-221``).  One kernel reads logits + labels once and emits the 3*N*C Dice sums;
the per-(sample, class) epilogue (a few dozen numbers) is ordinary differentiable PyTorch; the
backward kernel re-reads logits + labels and writes dlogits.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops

STRUCTURES = [  # reference capstone/utils/miccai.py:14-24; class id = index + 1
    "BrainStem", "Chiasm", "Mandible", "OpticNerve_L", "OpticNerve_R",
    "Parotid_L", "Parotid_R", "Submandibular_L", "Submandibular_R",
]
N_CLASSES = len(STRUCTURES) + 1


def _as_cl(input: torch.Tensor) -> torch.Tensor:
    """Logical (B, C, *S) logits -> channels-last (B, D, H, W, C); a view for UNet outputs."""
    if not input.is_cuda:
        raise RuntimeError("b200seg losses run on CUDA tensors only (no CPU fallback)")
    if input.dtype not in (torch.float32, torch.bfloat16):
        input = input.float()
    x = input.unsqueeze(2) if input.dim() == 4 else input
    if x.dim() != 5:
        raise ValueError(f"expected (B, C, H, W[, D]) logits, got {tuple(input.shape)}")
    cl = x.permute(0, 2, 3, 4, 1)
    try:
        ops.cl_info(cl)  # any uniform-stride channels-last layout (incl. padded buffers) is used as is
        return cl
    except ValueError:
        return cl.contiguous()


class _SoftmaxDiceSums(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits_cl: torch.Tensor, labels: torch.Tensor):
        ctx.save_for_backward(logits_cl, labels)
        return ops.softmax_dice_sums(logits_cl, labels)

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        logits_cl, labels = ctx.saved_tensors
        return ops.softmax_dice_bwd(logits_cl, labels, g[..., 0], g[..., 2]), None


class _FusedDiceLoss(torch.autograd.Function):
    """mean / sum softmax-Dice loss as three launches forward (sums, their reduction, the (B, C)
    epilogue with the gradient coefficients) and one backward kernel (+ a scalar scale)."""

    @staticmethod
    def forward(ctx, logits_cl, labels, include_background, smooth, mean):
        sums = ops.softmax_dice_sums(logits_cl, labels)
        loss, g_i, g_p = ops.dice_loss_epilogue(sums, include_background, smooth, mean)
        ctx.save_for_backward(logits_cl, labels, g_i, g_p)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits_cl, labels, g_i, g_p = ctx.saved_tensors
        return ops.softmax_dice_bwd(logits_cl, labels, g_i * g, g_p * g), None, None, None, None


def softmax_dice_sums(input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """(B, C, 3) = [I, G, P] with autograd through the fused kernels."""
    cl = _as_cl(input)
    if target.dim() == cl.dim() and target.shape[1] == 1:  # (B, 1, *S) as MONAI takes it
        target = target[:, 0]
    return _SoftmaxDiceSums.apply(cl, target)


class DiceLoss(nn.Module):
    """``DiceLoss(include_background, to_onehot_y, softmax, reduction)(input, target)`` with
    ``input`` (B, C, *S) logits and ``target`` (B, 1, *S) integer labels.  Only the
    configuration the reference uses is implemented: ``softmax=True, to_onehot_y=True``."""

    def __init__(self, include_background: bool = True, to_onehot_y: bool = False,
                 softmax: bool = False, reduction: str = "mean", smooth: float = 1e-5, **kwargs):
        super().__init__()
        if not (softmax and to_onehot_y):
            raise NotImplementedError("b200seg DiceLoss implements softmax=True, to_onehot_y=True")
        if kwargs.get("sigmoid") or kwargs.get("squared_pred") or kwargs.get("jaccard"):
            raise NotImplementedError("sigmoid / squared_pred / jaccard are not implemented")
        if reduction not in ("mean", "sum", "none"):
            raise ValueError(f"unsupported reduction {reduction}")
        self.include_background = include_background
        self.reduction = reduction
        self.smooth = float(smooth)

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.shape[1] != 1:
            raise AssertionError("labels must have a singleton channel dim (to_onehot_y=True)")
        if self.reduction in ("mean", "sum"):
            cl = _as_cl(input)
            return _FusedDiceLoss.apply(cl, target[:, 0], self.include_background, self.smooth,
                                        self.reduction == "mean")
        sums = softmax_dice_sums(input, target)
        if not self.include_background:
            sums = sums[:, 1:]
        inter, ground, pred = sums[..., 0], sums[..., 1], sums[..., 2]
        f = 1.0 - (2.0 * inter + self.smooth) / (ground + pred + self.smooth)
        if self.reduction == "mean":
            return f.mean()
        if self.reduction == "sum":
            return f.sum()
        return f


class GeneralizedDiceLoss(nn.Module):
    """In-tree ``GeneralizedDiceLoss`` of the reference (``capstone/models/temp.py:17-170``) on the
    same fused sums (softmax=True, to_onehot_y=True, batch=False)."""

    def __init__(self, include_background=True, to_onehot_y=False, softmax=False, w_type="square",
                 reduction="mean", smooth_nr=1e-5, smooth_dr=1e-5):
        super().__init__()
        if not (softmax and to_onehot_y):
            raise NotImplementedError("b200seg GeneralizedDiceLoss implements softmax=True, to_onehot_y=True")
        self.include_background, self.w_type, self.reduction = include_background, str(w_type), reduction
        self.smooth_nr, self.smooth_dr = float(smooth_nr), float(smooth_dr)

    def forward(self, input, target):
        sums = softmax_dice_sums(input, target)
        if not self.include_background:
            sums = sums[:, 1:]
        inter, ground, pred = sums[..., 0], sums[..., 1], sums[..., 2]
        g = ground.detach().float()
        if self.w_type == "simple":
            w = torch.reciprocal(g)
        elif self.w_type == "square":
            w = torch.reciprocal(g * g)
        else:
            w = torch.ones_like(g)
        infs = torch.isinf(w)
        w = torch.where(infs, torch.zeros_like(w), w)
        w = torch.where(infs, w.max(dim=1, keepdim=True).values.expand_as(w), w)
        f = 1.0 - (2.0 * (inter * w) + self.smooth_nr) / ((ground + pred) * w + self.smooth_dr)
        if self.reduction == "mean":
            return f.mean()
        if self.reduction == "sum":
            return f.sum()
        return f


# ---- reference wrapper API ------------------------------------------------------------------
class BaseLossWrapper(nn.Module):
    """``forward(input, target)`` with target (N, *S); adds the channel dim the loss expects
    (reference ``capstone/models/losses.py:24-42``; the 3-D variant drops the ndim assert,
    ``capstone/volumetric/losses.py:24-34``)."""

    def forward(self, input, target):
        return self.loss_fx(input, target.unsqueeze(dim=1))


class DiceLossWrapper(BaseLossWrapper):
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction
        self.loss_fx = DiceLoss(include_background=False, to_onehot_y=True, softmax=True,
                                reduction=reduction)


class GeneralizedDiceLossWrapper(BaseLossWrapper):
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction
        self.loss_fx = GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True,
                                           reduction=reduction)


DiceLossWrapper3D = DiceLossWrapper
GeneralizedDiceLossWrapper3D = GeneralizedDiceLossWrapper

LOSSES = {"Dice": DiceLossWrapper, "GeneralizedDice": GeneralizedDiceLossWrapper}


def apply_missing_mask(name, loss, mask_indicator):
    """Reference ``capstone/models/losses.py:206-221`` on the (N, C) loss matrix."""
    if name == "Focal":
        background = (mask_indicator.sum(dim=1, keepdim=True) == (N_CLASSES - 1)).float()
        mask_indicator = torch.cat([background, mask_indicator], dim=1)
    weights = 1.0 / mask_indicator.sum(dim=0)
    if torch.any(torch.isinf(weights)):
        weights = torch.ones_like(weights)
    weights = weights / weights.sum()
    return (loss * weights.unsqueeze(0) * mask_indicator).sum(dim=1).mean()


class MultipleLossWrapper(nn.Module):
    """``MultipleLossWrapper(losses, exclude_missing)(input, target, mask_indicator)`` ->
    dict of named losses (reference ``capstone/models/losses.py:170-203``;
    ``MultipleLossWrapper3D`` at ``capstone/volumetric/losses.py:128-130`` is the same with the
    ndim asserts removed -- SURVEY.md F7)."""

    def __init__(self, losses, exclude_missing=False):
        super().__init__()
        self.exclude_missing = exclude_missing
        for name in losses:
            if name not in LOSSES:
                raise NotImplementedError(
                    f"loss {name!r} is outside the B200 hot path (implemented: {sorted(LOSSES)})")
        reduction = "none" if exclude_missing else "mean"
        self.losses = nn.ModuleDict({name: LOSSES[name](reduction=reduction) for name in losses})

    def forward(self, input, target, mask_indicator=None, dist_maps=None):
        values = {}
        if mask_indicator is not None:
            mask_indicator = mask_indicator.float()
        for name, fx in self.losses.items():
            loss = fx(input, target)
            if self.exclude_missing:
                loss = apply_missing_mask(name, loss, mask_indicator)
            values[name] = loss
        return values


MultipleLossWrapper3D = MultipleLossWrapper
