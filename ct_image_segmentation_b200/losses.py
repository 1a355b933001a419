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
    def forward(ctx, logits_cl, labels, include_background, smooth, mean, counts_out=None):
        if counts_out is not None:  # the Dice METRIC counts of the same step, from the same pass over the logits
            sums, counts = ops.softmax_dice_metric_sums(logits_cl, labels)
            counts_out.append(counts)
        else:
            sums = ops.softmax_dice_sums(logits_cl, labels)
        loss, g_i, g_p = ops.dice_loss_epilogue(sums, include_background, smooth, mean)
        ctx.save_for_backward(logits_cl, labels, g_i, g_p)
        return loss

    @staticmethod
    def backward(ctx, g):
        logits_cl, labels, g_i, g_p = ctx.saved_tensors
        return ops.softmax_dice_bwd(logits_cl, labels, g_i * g, g_p * g), None, None, None, None, None


class _SoftmaxLossSums(torch.autograd.Function):
    """(B, C, 5) = [I, G, P, F, N] of ``b200seg_softmax_loss_fwd`` with autograd through the fused backward."""

    @staticmethod
    def forward(ctx, logits_cl, labels, gamma):
        ctx.save_for_backward(logits_cl, labels)
        ctx.gamma = gamma
        return ops.softmax_loss_sums(logits_cl, labels, gamma)

    @staticmethod
    def backward(ctx, g):
        logits_cl, labels = ctx.saved_tensors
        return ops.softmax_loss_bwd(logits_cl, labels, ctx.gamma, g[..., 0], g[..., 2], g[..., 3], g[..., 4]), None, None


class _SoftmaxBoundaryLossSums(torch.autograd.Function):
    """(B, C, 6) = [I, G, P, F, N, B] of ``b200seg_softmax_boundary_loss_fwd``: the shared pass with the Boundary
    loss's ``sum p_c * dist_{c-1}`` as a sixth accumulated sum."""

    @staticmethod
    def forward(ctx, logits_cl, labels, dist_maps, gamma):
        ctx.save_for_backward(logits_cl, labels, dist_maps)
        ctx.gamma = gamma
        return ops.softmax_boundary_loss_sums(logits_cl, labels, dist_maps, gamma)

    @staticmethod
    def backward(ctx, g):
        logits_cl, labels, dist_maps = ctx.saved_tensors
        return ops.softmax_boundary_loss_bwd(logits_cl, labels, dist_maps, ctx.gamma, g[..., 0], g[..., 2],
                                             g[..., 3], g[..., 4], g[..., 5]), None, None, None


def softmax_loss_sums(input: torch.Tensor, target: torch.Tensor, gamma: float = 2.0,
                      dist_maps: torch.Tensor = None) -> torch.Tensor:
    """One softmax pass for every voxel-wise loss of the reference: (B, C, 5) = [I, G, P, F, N]; with
    ``dist_maps`` (B, C-1, *S) a sixth column B = sum p_c * dist_{c-1} (Boundary loss)."""
    cl = _as_cl(input)
    if target.dim() == cl.dim() and target.shape[1] == 1:
        target = target[:, 0]
    if dist_maps is not None:
        return _SoftmaxBoundaryLossSums.apply(cl, target, dist_maps, float(gamma))
    return _SoftmaxLossSums.apply(cl, target, float(gamma))


def _n_voxels(input: torch.Tensor) -> int:
    n = 1
    for e in input.shape[2:]:
        n *= int(e)
    return n


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
                 softmax: bool = False, reduction: str = "mean", smooth: float = 1e-5, with_metric: bool = False,
                 **kwargs):
        """``with_metric`` (build-specific): also produce the Dice-metric counts {tp, |pred|, |target|} per (sample,
        class) of the same logits/labels in the SAME kernel pass (``self.metric_counts``, (B, C, 3) int64; feed
        ``metrics.dice_from_counts``) -- the reference evaluates that metric on every training step."""
        super().__init__()
        self.with_metric = bool(with_metric)
        self.metric_counts = None
        if not (softmax and to_onehot_y):
            raise NotImplementedError("b200seg DiceLoss implements softmax=True, to_onehot_y=True")
        if kwargs.get("sigmoid") or kwargs.get("squared_pred") or kwargs.get("jaccard"):
            raise NotImplementedError("sigmoid / squared_pred / jaccard are not implemented")
        if reduction not in ("mean", "sum", "none"):
            raise ValueError(f"unsupported reduction {reduction}")
        self.include_background = include_background
        self.reduction = reduction
        self.smooth = float(smooth)

    def forward(self, input: torch.Tensor, target: torch.Tensor, sums=None) -> torch.Tensor:
        if target.shape[1] != 1:
            raise AssertionError("labels must have a singleton channel dim (to_onehot_y=True)")
        if sums is None and self.reduction in ("mean", "sum"):
            cl = _as_cl(input)
            box = [] if self.with_metric else None
            loss = _FusedDiceLoss.apply(cl, target[:, 0], self.include_background, self.smooth,
                                        self.reduction == "mean", box)
            if box:
                self.metric_counts = box[0]
            return loss
        if sums is None:
            if self.with_metric:
                with torch.no_grad():
                    _, self.metric_counts = ops.softmax_dice_metric_sums(_as_cl(input.detach()), target[:, 0])
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

    def forward(self, input, target, sums=None):
        if sums is None:
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


class FocalLoss(nn.Module):
    """``monai.losses.FocalLoss(gamma=2.0, reduction)`` on a one-hot target, as the reference builds it
    (``capstone/models/losses.py:105-124``): per (sample, class) the mean over voxels of
    ``-(1 - p)^gamma * t * log p`` -> (B, C); ``mean`` / ``sum`` reduce that matrix.  ``target`` is the
    label map (B, 1, *S) the one-hot is made from."""

    def __init__(self, gamma: float = 2.0, reduction: str = "mean"):
        super().__init__()
        self.gamma, self.reduction = float(gamma), reduction

    def forward(self, input, target, sums=None):
        if sums is None:
            sums = softmax_loss_sums(input, target, self.gamma)
        f = sums[..., 3] / float(_n_voxels(input))
        if self.reduction == "mean":
            return f.mean()
        if self.reduction == "sum":
            return f.sum()
        return f


class BoundaryLoss(nn.Module):
    """The reference's ``BoundaryLossWrapper`` arithmetic (``capstone/models/losses.py:127-157``):
    ``softmax(input, 1)[:, 1:] * dist_maps`` averaged over everything (``mean``) or over the spatial axes
    (``none`` -> (B, C-1)); ``dist_maps`` (B, C-1, *S) are the dataset's pre-computed signed distance maps
    (``capstone/data/utils.py:10-26``).  Needs no labels; when it shares the softmax pass with other losses the
    sums come in through ``sums`` (column 5)."""

    def __init__(self, reduction: str = "mean"):
        super().__init__()
        if reduction not in ("none", "mean"):
            raise AssertionError("reduction must be 'none' or 'mean'")
        self.reduction = reduction

    def forward(self, input, dist_maps, sums=None):
        if sums is None:
            b = input.shape[0]
            dummy = torch.zeros((b, 1) + tuple(input.shape[2:]), dtype=torch.uint8, device=input.device)
            sums = softmax_loss_sums(input, dummy, dist_maps=dist_maps)
        per = sums[:, 1:, 5] / float(_n_voxels(input))  # (B, C-1): spatial mean per sample and class
        return per if self.reduction == "none" else per.mean()


class CrossEntropyLoss(nn.Module):
    """``F.cross_entropy(input, target[, weight])`` (mean reduction) as the reference's
    ``CrossEntropyWrapper`` / ``WeightedCrossEntropyWrapper`` call it (``capstone/models/losses.py:45-68``)."""

    def __init__(self, weight=None):
        super().__init__()
        self.register_buffer("weight", None if weight is None else torch.as_tensor(weight, dtype=torch.float32))

    def forward(self, input, target, sums=None):
        if sums is None:
            sums = softmax_loss_sums(input, target)
        nll, cnt = sums[..., 4], sums[..., 1]
        if self.weight is None:
            return nll.sum() / cnt.sum()
        w = self.weight.to(nll.device)
        return (nll * w).sum() / (cnt * w).sum()


WEIGHT = {  # inverse pixel frequency, reference capstone/models/losses.py:10-21
    "Background": 1e-10, "BrainStem": 0.007, "Chiasm": 0.3296, "Mandible": 0.0046, "OpticNerve_L": 0.2619,
    "OpticNerve_R": 0.3035, "Parotid_L": 0.0068, "Parotid_R": 0.0065, "Submandibular_L": 0.0374,
    "Submandibular_R": 0.0426,
}


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


class FocalLossWrapper(BaseLossWrapper):
    def __init__(self, reduction="mean"):
        super().__init__()
        self.reduction = reduction
        self.loss_fx = FocalLoss(reduction=reduction)


class CrossEntropyWrapper(nn.Module):
    """Takes (and ignores) the ``reduction`` keyword like the reference (``**kwargs``): always the mean."""

    def __init__(self, **kwargs):
        super().__init__()
        self.loss_fx = CrossEntropyLoss()

    def forward(self, input, target):
        return self.loss_fx(input, target)


class WeightedCrossEntropyWrapper(CrossEntropyWrapper):
    def __init__(self, **kwargs):
        super().__init__()
        self.loss_fx = CrossEntropyLoss(weight=list(WEIGHT.values()))


class BoundaryLossWrapper(nn.Module):
    """``BoundaryLossWrapper(reduction)(input, dist_maps)`` (reference ``capstone/models/losses.py:127-157``; the
    2-D ndim asserts are dropped as in the 3-D twins of the other wrappers)."""

    def __init__(self, reduction="mean"):
        super().__init__()
        assert reduction in ["none", "mean"]
        self.reduction = reduction
        self.loss_fx = BoundaryLoss(reduction=reduction)

    def forward(self, input, dist_maps):
        return self.loss_fx(input, dist_maps)


DiceLossWrapper3D = DiceLossWrapper
GeneralizedDiceLossWrapper3D = GeneralizedDiceLossWrapper
FocalLossWrapper3D = FocalLossWrapper
CrossEntropyWrapper3D = CrossEntropyWrapper
WeightedCrossEntropyWrapper3D = WeightedCrossEntropyWrapper

LOSSES = {"CrossEntropy": CrossEntropyWrapper, "WeightedCrossEntropy": WeightedCrossEntropyWrapper,
          "Focal": FocalLossWrapper, "Dice": DiceLossWrapper, "GeneralizedDice": GeneralizedDiceLossWrapper,
          "Boundary": BoundaryLossWrapper}


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

    def enable_metric_counts(self, on: bool = True) -> bool:
        """Ask the Dice loss (if it is one of the losses) to produce the Dice-METRIC counts of the same logits in its
        own kernel pass; ``metric_counts`` then holds them after every ``forward``.  Returns whether it applies."""
        fx = self.losses["Dice"].loss_fx if "Dice" in self.losses else None
        if fx is None:
            return False
        fx.with_metric = bool(on)
        fx.metric_counts = None
        return True

    @property
    def metric_counts(self):
        fx = self.losses["Dice"].loss_fx if "Dice" in self.losses else None
        return None if fx is None else fx.metric_counts

    def forward(self, input, target, mask_indicator=None, dist_maps=None):
        values = {}
        if mask_indicator is not None:
            mask_indicator = mask_indicator.float()
        if "Boundary" in self.losses:
            assert dist_maps is not None, "Distance maps are required for using boundary loss"
        # Focal / CrossEntropy / Boundary present: ONE softmax pass feeds every requested loss (the reference runs a
        # softmax / log_softmax per loss); Dice alone keeps its 3-sum kernel and one-launch epilogue
        shared = None
        if any(n in ("Focal", "CrossEntropy", "WeightedCrossEntropy", "Boundary") for n in self.losses):
            shared = softmax_loss_sums(input, target.unsqueeze(1),
                                       dist_maps=dist_maps if "Boundary" in self.losses else None)
        for name, fx in self.losses.items():
            if shared is None:
                loss = fx(input, target)
            elif name == "Boundary":
                loss = fx.loss_fx(input, dist_maps, sums=shared)
            elif name in ("CrossEntropy", "WeightedCrossEntropy"):
                loss = fx.loss_fx(input, target, sums=shared)
            else:
                loss = fx.loss_fx(input, target.unsqueeze(1), sums=shared)
            if self.exclude_missing and name not in ("CrossEntropy", "WeightedCrossEntropy"):
                loss = apply_missing_mask(name, loss, mask_indicator)
            values[name] = loss
        return values


MultipleLossWrapper3D = MultipleLossWrapper
