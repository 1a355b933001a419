"""The CPU oracle against the reference's own known answers (no GPU needed).

Known answers: parameter counts from reports/Report.pdf Table 1 (exact values SURVEY.md A.5),
the module path used by capstone/interpretability.py:88, and fixtures produced by running the
reference's in-tree functions (tests/golden/make_golden.py).
"""
import numpy as np
import pytest
import torch

from oracle import monai_ref as O

PARAMS = [
    (2, [64, 128, 256, 512, 1024], 1, (10557650, 13408292, 25980905)),
    (2, [64, 128, 256, 512, 1024], 3, (10558802, 13410596, 25983209)),
    (3, [16, 32, 64, 128, 256], 1, (1986530, 2458508, 4816001)),
    (3, [32, 64, 128, 256, 512], 1, (7926706, 9804396, 19233361)),
]


@pytest.mark.parametrize("dim,ch,inc,expected", PARAMS)
def test_param_counts(dim, ch, inc, expected):
    for r in (0, 1, 2):
        net = O.UNet(dim, inc, 10, ch, [2, 2, 2, 2], num_res_units=r)
        assert O.count_parameters(net) == expected[r]


def test_module_path_and_keys():
    net = O.UNet(3, 1, 10, [16, 32, 64, 128, 256], [2, 2, 2, 2], num_res_units=2)
    assert isinstance(net.model[2][1].conv.unit0.conv, torch.nn.Conv3d)  # interpretability.py:88
    sd = net.state_dict()
    assert len(sd) == 63
    assert tuple(sd["model.2.0.conv.weight"].shape) == (32, 10, 3, 3, 3)  # ConvTranspose layout
    bottom = "model.1.submodule.1.submodule.1.submodule.1.submodule."
    assert tuple(sd[bottom + "residual.weight"].shape) == (256, 128, 1, 1, 1)
    assert "model.2.1.conv.unit0.act.weight" not in sd  # conv-only head


def test_forward_shapes():
    net = O.UNet(3, 1, 10, [4, 8, 8, 16, 16], [2, 2, 2, 2], num_res_units=2)
    y = net(torch.randn(1, 1, 16, 32, 16))
    assert tuple(y.shape) == (1, 10, 16, 32, 16)
    net2 = O.UNet(2, 3, 10, [4, 8, 8, 16, 16], [2, 2, 2, 2], num_res_units=0)
    assert tuple(net2(torch.randn(2, 3, 32, 48)).shape) == (2, 10, 32, 48)


@pytest.mark.parametrize("tag", ["dense", "sparse"])
def test_dice_loss_vs_reference_intree(golden, tag):
    logits = torch.from_numpy(golden["dice_logits"]).requires_grad_(True)
    lab = torch.from_numpy(golden[f"dice_lab_{tag}"]).long().unsqueeze(1)
    fx = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="mean")
    v = fx(logits, lab)
    v.backward()
    np.testing.assert_allclose(v.item(), golden[f"dice_{tag}_mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(logits.grad.numpy(), golden[f"dice_{tag}_grad"], rtol=1e-5, atol=1e-9)
    fxn = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="none")
    np.testing.assert_allclose(fxn(logits.detach(), lab).numpy(), golden[f"dice_{tag}_none"],
                               rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("wt", ["square", "simple"])
@pytest.mark.parametrize("tag", ["dense", "sparse"])
def test_generalized_dice_vs_reference_intree(golden, wt, tag):
    """capstone/models/temp.py GeneralizedDiceLoss, its default weighting incl. inf -> max (sparse: absent classes)."""
    logits = torch.from_numpy(golden["dice_logits"]).requires_grad_(True)
    lab = torch.from_numpy(golden[f"dice_lab_{tag}"]).long().unsqueeze(1)
    v = O.GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True, w_type=wt)(logits, lab)
    v.backward()
    np.testing.assert_allclose(v.item(), golden[f"gdl_{wt}_{tag}_mean"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(logits.grad.numpy(), golden[f"gdl_{wt}_{tag}_grad"], rtol=1e-5, atol=1e-9)
    fxn = O.GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True, w_type=wt, reduction="none")
    np.testing.assert_allclose(fxn(logits.detach(), lab).numpy(), golden[f"gdl_{wt}_{tag}_none"], rtol=1e-6, atol=1e-7)


def test_boundary_loss_vs_reference(golden):
    """capstone/models/losses.py BoundaryLossWrapper / MultipleLossWrapper(["Boundary"], exclude_missing=True)."""
    logits = torch.from_numpy(golden["boundary_logits"]).requires_grad_(True)
    dist = torch.from_numpy(golden["boundary_dist"])
    v = O.BoundaryLoss("mean")(logits, dist)
    v.backward()
    np.testing.assert_allclose(v.item(), golden["boundary_mean"], rtol=1e-6)
    np.testing.assert_allclose(logits.grad.numpy(), golden["boundary_grad"], rtol=1e-5, atol=1e-10)
    np.testing.assert_allclose(O.BoundaryLoss("none")(logits.detach(), dist).numpy(), golden["boundary_none"], rtol=1e-6)
    out = O.MultipleLossWrapper(["Boundary"], exclude_missing=True)(
        logits.detach(), torch.from_numpy(golden["boundary_lab"]).long(), torch.from_numpy(golden["indicator"]), dist)
    np.testing.assert_allclose(out["Boundary"].item(), golden["boundary_missing"], rtol=1e-6)


def test_dice_loss_2d(golden):
    fx = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="none")
    v = fx(torch.from_numpy(golden["dice2d_logits"]),
           torch.from_numpy(golden["dice2d_lab"]).long().unsqueeze(1))
    np.testing.assert_allclose(v.numpy(), golden["dice2d_none"], rtol=1e-6, atol=1e-7)


def test_missing_mask(golden):
    per = torch.from_numpy(golden["dice_sparse_none"])
    for ind_key, out_key in (("indicator", "missing_dice"), ("indicator_inf", "missing_dice_inf")):
        v = O.apply_missing_mask("Dice", per, torch.from_numpy(golden[ind_key]))
        np.testing.assert_allclose(v.numpy(), golden[out_key], rtol=1e-6)
    v = O.apply_missing_mask("Focal", torch.from_numpy(golden["focal_like"]),
                             torch.from_numpy(golden["indicator"]))
    np.testing.assert_allclose(v.numpy(), golden["missing_focal"], rtol=1e-6)


def test_label_maps(golden):
    m3 = torch.from_numpy(golden["masks"])
    assert np.array_equal(O.squash_masks(m3).numpy(), golden["squash3d"])
    m2 = torch.from_numpy(golden["masks2d"])
    assert np.array_equal(O.squash_masks(m2).numpy(), golden["squash2d"])
    assert np.array_equal(O.squash_predictions(torch.from_numpy(golden["dice_logits"])).numpy(),
                          golden["argmax"])
    assert np.array_equal(O.squash_predictions(torch.from_numpy(golden["tie_logits"])).numpy(),
                          golden["tie_argmax"])


def test_dice_metric(golden):
    pred = torch.from_numpy(golden["argmax"])
    for tag in ("dense", "sparse"):
        dm, dpc = O.dice_metric(pred, torch.from_numpy(golden[f"dice_lab_{tag}"]))
        np.testing.assert_allclose(dpc.numpy(), golden[f"metric_{tag}_per_class"], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(dm.item(), golden[f"metric_{tag}_mean"], rtol=1e-6, atol=1e-7)
    dm, dpc = O.dice_metric(torch.from_numpy(golden["metric_noisy_pred"]),
                            torch.from_numpy(golden["metric_noisy_target"]))
    np.testing.assert_allclose(dpc.numpy(), golden["metric_noisy_per_class"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dm.item(), golden["metric_noisy_mean"], rtol=1e-6)
    assert dpc[3].item() == 0.0  # class 4 absent from every target -> NaN -> 0 after reduction


def test_windowing(golden):
    hu = golden["hu"]
    for name in O.WINDOWING_CONFIG:
        w, l = golden[f"window_{name}_cfg"]
        assert (int(w), int(l)) == O.WINDOWING_CONFIG[name]
        np.testing.assert_array_equal(O.apply_window(hu, int(w), int(l)), golden[f"window_{name}"])
    np.testing.assert_array_equal(O.apply_window(hu, 350, 20, shift=False), golden["window_soft_noshift"])
    # soft tissue clips to [-155, 195]
    assert golden["window_soft_noshift"].min() == -155 and golden["window_soft_noshift"].max() == 195


def test_sliding_window_identity():
    x = torch.randn(1, 1, 40, 36, 20)
    pred = lambda b: torch.cat([b, 2 * b], 1)
    y = O.sliding_window_inference(x, (16, 16, 16), 4, pred, overlap=0.25)
    torch.testing.assert_close(y, torch.cat([x, 2 * x], 1), rtol=1e-5, atol=1e-6)
    assert O._scan_starts(512, 128, 0.25) == [0, 96, 192, 288, 384]
    assert O._scan_starts(160, 128, 0.25) == [0, 32]
