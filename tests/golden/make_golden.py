"""Generate golden fixtures by running the REFERENCE's own in-tree functions.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python tests/golden/make_golden.py``.  Writes ``tests/golden/*.npz``.

The reference's hot-path files import third-party packages that are not
installed (monai, albumentations, nrrd, pytorch_lightning).  Only the few
symbols those files need at import time are stubbed below; every function whose
OUTPUT is recorded is the reference's own code, executed unmodified from
``/root/reference``:

* ``capstone/models/temp.py``: ``GeneralizedDiceLoss`` (with ``w_type="uniform"``
  it is exactly the MONAI-0.3 DiceLoss formula), ``compute_meandice``,
  ``do_metric_reduction``
* ``capstone/models/losses.py``: ``apply_missing_mask``, ``BoundaryLossWrapper``, ``MultipleLossWrapper``
  (with the Boundary loss, the only entry of its table that needs no monai class)
* ``capstone/models/metrics.py``: ``DiceMetricWrapper`` (via the 3D subclass)
* ``capstone/volumetric/utils.py``: ``_squash_masks_3D``;
  ``capstone/training/utils.py``: ``_squash_masks``, ``_squash_predictions``
* ``capstone/transforms/transforms_2d.py``: ``apply_window``, ``WINDOWING_CONFIG``
"""
import enum
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install_stubs():
    def one_hot(labels, num_classes, dtype=torch.float, dim=1):
        sh = list(labels.shape)
        sh[dim] = num_classes
        o = torch.zeros(size=sh, dtype=dtype, device=labels.device)
        return o.scatter_(dim=dim, index=labels.long(), value=1)

    class LossReduction(enum.Enum):
        NONE = "none"
        MEAN = "mean"
        SUM = "sum"

    class MetricReduction(enum.Enum):
        NONE = "none"
        MEAN = "mean"
        SUM = "sum"
        MEAN_BATCH = "mean_batch"
        SUM_BATCH = "sum_batch"
        MEAN_CHANNEL = "mean_channel"
        SUM_CHANNEL = "sum_channel"

    class Weight(enum.Enum):
        SQUARE = "square"
        SIMPLE = "simple"
        UNIFORM = "uniform"

    class AsDiscrete:
        def __init__(self, to_onehot=False, n_classes=None, **kw):
            self.to_onehot, self.n_classes = to_onehot, n_classes

        def __call__(self, x):
            return one_hot(x, self.n_classes) if self.to_onehot else x

    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError("third-party symbol not available (stub)")

    _stub("monai")
    _stub("monai.networks", one_hot=one_hot)
    _stub("monai.networks.nets", UNet=_Missing)
    _stub("monai.utils", LossReduction=LossReduction, MetricReduction=MetricReduction, Weight=Weight)
    _stub("monai.losses")
    _stub("monai.losses.dice", DiceLoss=_Missing)
    _stub("monai.losses.focal_loss", FocalLoss=_Missing)
    _stub("monai.transforms", AsDiscrete=AsDiscrete)
    _stub("nrrd")
    _stub("albumentations")
    _stub("albumentations.core")
    _stub("albumentations.core.transforms_interface", ImageOnlyTransform=object)
    # capstone/utils/__init__ pulls visualize -> matplotlib etc.; provide the constant only
    pkg = _stub("capstone")
    pkg.__path__ = [os.path.join(REF, "capstone")]
    utils = _stub("capstone.utils")
    utils.__path__ = []
    structures = [
        "BrainStem", "Chiasm", "Mandible", "OpticNerve_L", "OpticNerve_R",
        "Parotid_L", "Parotid_R", "Submandibular_L", "Submandibular_R",
    ]
    miccai = _stub("capstone.utils.miccai", STRUCTURES=structures)
    utils.miccai = miccai
    # capstone/training/__init__ imports the LightningModules; bypass it
    tr = _stub("capstone.training")
    tr.__path__ = [os.path.join(REF, "capstone", "training")]
    tf = _stub("capstone.transforms")
    tf.__path__ = [os.path.join(REF, "capstone", "transforms")]
    vol = _stub("capstone.volumetric")
    vol.__path__ = [os.path.join(REF, "capstone", "volumetric")]


def ellipsoid_masks(gen, b, shape, n_struct=9, drop=()):
    """(b, 9, *shape) uint8 masks of axis-aligned ellipsoids (sparse foreground)."""
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    m = np.zeros((b, n_struct) + tuple(shape), dtype=np.uint8)
    for i in range(b):
        for c in range(n_struct):
            if (i, c) in drop:
                continue
            ctr = [gen.uniform(0.2, 0.8) * s for s in shape]
            rad = [max(1.5, gen.uniform(0.06, 0.16) * s) for s in shape]
            d = sum(((g - c0) / r) ** 2 for g, c0, r in zip(grids, ctr, rad))
            m[i, c] = (d <= 1.0).astype(np.uint8)
    return m


def main():
    assert os.path.isdir(REF), "run in the build container (needs /root/reference)"
    install_stubs()
    # structure constant is taken verbatim from the reference file (cannot import it: nrrd)
    src = open(os.path.join(REF, "capstone/utils/miccai.py")).read()
    for s in sys.modules["capstone.utils.miccai"].STRUCTURES:
        assert f'"{s}"' in src

    from capstone.models import temp as ref_temp
    from capstone.models.losses import apply_missing_mask as ref_apply_missing_mask
    from capstone.volumetric.metrics import DiceMetricWrapper3D
    from capstone.volumetric.utils import _squash_masks_3D
    from capstone.training.utils import _squash_masks, _squash_predictions
    from capstone.transforms.transforms_2d import WINDOWING_CONFIG, apply_window

    gen = np.random.default_rng(12342)
    torch.manual_seed(12342)
    out = {}

    # --- Dice loss (3D, 10 classes): dense-random and sparse labels, mean + none -------
    B, C, S = 2, 10, (8, 12, 10)
    logits = torch.randn(B, C, *S) * 2.0
    lab_dense = torch.randint(0, C, (B, *S))
    masks = torch.from_numpy(ellipsoid_masks(gen, B, S, drop={(1, 3), (1, 7)}))
    lab_sparse = _squash_masks_3D(masks, C, "cpu")
    out["dice_logits"] = logits.numpy()
    out["dice_lab_dense"] = lab_dense.numpy().astype(np.uint8)
    out["masks"] = masks.numpy()
    out["dice_lab_sparse"] = lab_sparse.numpy().astype(np.uint8)
    for tag, lab in (("dense", lab_dense), ("sparse", lab_sparse)):
        for red in ("mean", "none"):
            fx = ref_temp.GeneralizedDiceLoss(include_background=False, to_onehot_y=True,
                                              softmax=True, w_type="uniform", reduction=red)
            lg = logits.clone().requires_grad_(True)
            v = fx(lg, lab.unsqueeze(1))
            out[f"dice_{tag}_{red}"] = v.detach().numpy()
            if red == "mean":
                v.backward()
                out[f"dice_{tag}_grad"] = lg.grad.numpy()
    # 2D case
    logits2 = torch.randn(3, C, 16, 20)
    lab2 = torch.randint(0, C, (3, 16, 20))
    fx = ref_temp.GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True,
                                      w_type="uniform", reduction="none")
    out["dice2d_logits"] = logits2.numpy()
    out["dice2d_lab"] = lab2.numpy().astype(np.uint8)
    out["dice2d_none"] = fx(logits2, lab2.unsqueeze(1)).numpy()

    # --- missing-annotation masking -----------------------------------------------------
    ind = torch.ones(B, 9)
    ind[1, 3] = 0
    ind[1, 7] = 0
    per = torch.from_numpy(out["dice_sparse_none"])
    out["indicator"] = ind.numpy()
    out["missing_dice"] = ref_apply_missing_mask("Dice", per.clone(), ind.clone()).numpy()
    ind0 = ind.clone()
    ind0[:, 2] = 0  # a class with no annotation in the batch -> inf -> ones path
    out["indicator_inf"] = ind0.numpy()
    out["missing_dice_inf"] = ref_apply_missing_mask("Dice", per.clone(), ind0.clone()).numpy()
    focal_like = torch.rand(B, 10)
    out["focal_like"] = focal_like.numpy()
    out["missing_focal"] = ref_apply_missing_mask("Focal", focal_like.clone(), ind.clone()).numpy()

    # --- label maps ----------------------------------------------------------------------
    out["squash3d"] = lab_sparse.numpy().astype(np.uint8)
    m2 = torch.from_numpy(ellipsoid_masks(gen, 2, (24, 20)))
    out["masks2d"] = m2.numpy()
    out["squash2d"] = _squash_masks(m2, C, "cpu").numpy().astype(np.uint8)
    out["argmax"] = _squash_predictions(logits).numpy().astype(np.uint8)
    # softmax-saturation ties: huge logits make softmax produce exact duplicates
    tie = torch.randn(1, C, 4, 4, 4)
    tie[:, 2] = 200.0
    tie[:, 5] = 200.0 + 1e-5   # same float after softmax saturation? keep as recorded
    tie[:, 7] = tie[:, 4]      # exact duplicate channels
    out["tie_logits"] = tie.numpy()
    out["tie_argmax"] = _squash_predictions(tie).numpy().astype(np.uint8)

    # --- Dice metric -----------------------------------------------------------------------
    pred_lab = _squash_predictions(logits)
    wrapper = DiceMetricWrapper3D()
    for tag, lab in (("dense", lab_dense), ("sparse", lab_sparse)):
        dm, dpc = wrapper(pred_lab.clone(), lab.clone())
        out[f"metric_{tag}_mean"] = dm.numpy()
        out[f"metric_{tag}_per_class"] = dpc.numpy()
    # a prediction that overlaps the sparse target a lot, with one class absent everywhere
    noisy = lab_sparse.clone()
    flip = torch.rand(noisy.shape) < 0.1
    noisy[flip] = torch.randint(0, C, (int(flip.sum()),))
    t_abs = lab_sparse.clone()
    t_abs[t_abs == 4] = 0  # class 4 has no ground truth in any sample -> NaN -> 0
    dm, dpc = wrapper(noisy.clone(), t_abs.clone())
    out["metric_noisy_pred"] = noisy.numpy().astype(np.uint8)
    out["metric_noisy_target"] = t_abs.numpy().astype(np.uint8)
    out["metric_noisy_mean"] = dm.numpy()
    out["metric_noisy_per_class"] = dpc.numpy()
    oh = lambda x: sys.modules["monai.networks"].one_hot(x.unsqueeze(1), C)
    out["meandice_raw"] = ref_temp.compute_meandice(oh(noisy), oh(t_abs),
                                                    include_background=False).numpy()

    # --- HU windowing -------------------------------------------------------------------------
    hu = gen.integers(-1024, 3072, size=(32, 40, 1)).astype(np.int16)
    out["hu"] = hu
    for name, (w, l) in WINDOWING_CONFIG.items():
        out[f"window_{name}"] = apply_window(hu, w, l)
        out[f"window_{name}_cfg"] = np.asarray([w, l])
    out["window_soft_noshift"] = apply_window(hu, *WINDOWING_CONFIG["soft_tissue"], shift=False)

    # --- round 2 additions (appended so that the RNG stream of everything above is unchanged) ---------------
    # GeneralizedDiceLoss with the reference's own default weighting ("square") and "simple", incl. the inf -> max
    # rule (capstone/models/temp.py:146-152): the sparse labels have classes with no voxel in sample 1
    for wt in ("square", "simple"):
        for tag, lab in (("dense", lab_dense), ("sparse", lab_sparse)):
            for red in ("mean", "none"):
                fx = ref_temp.GeneralizedDiceLoss(include_background=False, to_onehot_y=True,
                                                  softmax=True, w_type=wt, reduction=red)
                lg = logits.clone().requires_grad_(True)
                v = fx(lg, lab.unsqueeze(1))
                out[f"gdl_{wt}_{tag}_{red}"] = v.detach().numpy()
                if red == "mean":
                    v.backward()
                    out[f"gdl_{wt}_{tag}_grad"] = lg.grad.numpy()

    # Boundary loss (capstone/models/losses.py:127-157) on 2-D logits with signed distance maps.  The maps are
    # INPUTS: built here with scipy the way capstone/data/utils.py:10-26 does (that file itself uses np.bool,
    # removed from current numpy, and cannot run).
    from capstone.models.losses import BoundaryLossWrapper as RefBoundary
    from capstone.models.losses import MultipleLossWrapper as RefMultiple
    from scipy.ndimage import distance_transform_edt as edt
    mb = m2.numpy().astype(bool)                      # (2, 9, 24, 20)
    dist = np.zeros(mb.shape, dtype=np.float32)
    for i in range(mb.shape[0]):
        for c in range(mb.shape[1]):
            pos = mb[i, c]
            if pos.any():
                neg = ~pos
                dist[i, c] = edt(neg) * neg - (edt(pos) - 1) * pos
    dist /= 255.0
    dist_t = torch.from_numpy(dist)
    logits_b = torch.randn(2, C, 24, 20) * 2.0
    out["boundary_logits"] = logits_b.numpy()
    out["boundary_dist"] = dist
    for red in ("mean", "none"):
        lg = logits_b.clone().requires_grad_(True)
        v = RefBoundary(reduction=red)(lg, dist_t)
        out[f"boundary_{red}"] = v.detach().numpy()
        if red == "mean":
            v.backward()
            out["boundary_grad"] = lg.grad.numpy()
    lab_b = _squash_masks(m2, C, "cpu")
    out["boundary_lab"] = lab_b.numpy().astype(np.uint8)
    lg = logits_b.clone().requires_grad_(True)
    vals = RefMultiple(["Boundary"], exclude_missing=True)(lg, lab_b, ind.clone(), dist_t)
    out["boundary_missing"] = vals["Boundary"].detach().numpy()
    vals["Boundary"].backward()
    out["boundary_missing_grad"] = lg.grad.numpy()

    path = os.path.join(HERE, "ref_intree.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
