"""Sliding-window inference and the patch sampler against the oracle -- `pytest -m gpu`."""
import numpy as np
import pytest
import torch

import ct_image_segmentation_b200 as B
from ct_image_segmentation_b200.inference import scan_starts, sliding_window_inference, window_list
from ct_image_segmentation_b200.sampler import PatchSampler
from oracle import monai_ref as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_sliding_window_matches_oracle_fp32():
    torch.manual_seed(12342)
    ch = [8, 16, 16, 32, 32]
    ref = O.UNet(3, 1, 10, ch, [2, 2, 2, 2], num_res_units=2)
    net = B.UNet(3, 1, 10, ch, [2, 2, 2, 2], num_res_units=2, dtype=torch.float32)
    net.load_state_dict(ref.state_dict())
    net = net.to(DEV)
    x = torch.randn(1, 1, 40, 72, 56)
    roi = (32, 48, 32)
    with torch.no_grad():
        want = O.sliding_window_inference(x, roi, 2, ref, overlap=0.25)
    labels, logits = sliding_window_inference(x.to(DEV), roi, 2, net, overlap=0.25, return_logits=True)
    assert tuple(labels.shape) == (1, 40, 72, 56) and labels.dtype == torch.uint8
    err = ((logits.cpu() - want).norm() / want.norm()).item()
    assert err < 1e-4, err
    lm = O.squash_predictions(want)
    mism = labels.cpu().long() != lm
    if mism.any():  # only exact-tie voxels may differ (1e-6 logit noise)
        top2 = torch.softmax(want, 1).topk(2, dim=1).values
        assert float((top2[:, 0] - top2[:, 1])[mism].max()) < 1e-5


@pytest.mark.parametrize("mode", ["constant", "gaussian"])
@pytest.mark.parametrize("world", [2, 3, 8])
def test_sliding_window_sharded_is_bit_identical(world, mode):
    """The world-rank algorithm (windows in contiguous runs, output slabs owned by ranks, row runs of the bf16
    predictions handed to the slab owners, accumulation in global window order) emulated rank after rank on one GPU:
    the label map AND the averaged logits equal the single-rank run bit for bit, for every world size."""
    from ct_image_segmentation_b200.inference import emulate_ranks
    torch.manual_seed(1)
    net = B.UNet(3, 1, 10, [8, 16, 16], [2, 2], num_res_units=1, dtype=torch.bfloat16).to(DEV)
    x = torch.randn(1, 1, 40, 40, 24, device=DEV)
    roi = (16, 16, 16)
    full, full_logits = sliding_window_inference(x, roi, 4, net, 0.25, mode=mode, return_logits=True, rank=0, world=1)
    lab, logits, plan = emulate_ranks(x, roi, 4, net, world, 0.25, mode=mode)
    assert torch.equal(lab, full) and torch.equal(logits, full_logits)
    # geometry: every window is computed exactly once, its rows go to exactly the slabs they fall in
    assert plan.runs[0] == 0 and plan.runs[-1] == len(plan.wins) == 3 * 3 * 2
    assert sum(p.e - p.s for p in plan.pieces) == len(plan.wins) * roi[0]
    assert all(plan.bounds[p.dst] <= p.s < p.e <= plan.bounds[p.dst + 1] for p in plan.pieces)
    assert scan_starts(512, 128, 0.25) == [0, 96, 192, 288, 384] and scan_starts(160, 128, 0.25) == [0, 32]
    assert full.dtype == torch.uint8 and int(full.max()) <= 9


def test_sliding_window_sharded_bit_identical_on_sliding_kernels():
    """The same property at a ROI large enough for the sliding-window tcgen05 kernels and the deferred InstanceNorm
    statistics (their grids depend on the batch size): ragged per-rank batches are padded to the full batch size, so
    windows computed in another batch slot / on another rank are bit-identical and so is the label map."""
    from ct_image_segmentation_b200.inference import GraphedPredictor, emulate_ranks
    torch.manual_seed(3)
    net = B.UNet(3, 1, 10, [16, 32, 64], [2, 2], num_res_units=2, dtype=torch.bfloat16).to(DEV)
    x = torch.randn(1, 1, 80, 96, 64, device=DEV)
    roi = (32, 64, 64)
    pred = GraphedPredictor(net, torch.zeros(2, 1, *roi, device=DEV))
    full, full_logits = sliding_window_inference(x, roi, 2, pred, 0.25, return_logits=True, rank=0, world=1)
    eager = sliding_window_inference(x, roi, 2, net, 0.25, rank=0, world=1)
    assert torch.equal(eager, full)
    for world in (2, 4):
        lab, logits, plan = emulate_ranks(x, roi, 2, pred, world, 0.25)
        assert len(plan.wins) == 6
        assert torch.equal(logits, full_logits) and torch.equal(lab, full), world


def test_sliding_window_gaussian_matches_oracle():
    torch.manual_seed(7)
    ch = [8, 16, 16]
    ref = O.UNet(3, 1, 10, ch, [2, 2], num_res_units=1)
    net = B.UNet(3, 1, 10, ch, [2, 2], num_res_units=1, dtype=torch.float32)
    net.load_state_dict(ref.state_dict())
    net = net.to(DEV)
    x = torch.randn(1, 1, 24, 40, 28)
    roi = (16, 16, 16)
    with torch.no_grad():
        want = O.sliding_window_inference(x, roi, 3, ref, overlap=0.25, mode="gaussian")
        const = O.sliding_window_inference(x, roi, 3, ref, overlap=0.25)
    labels, logits = sliding_window_inference(x.to(DEV), roi, 3, net, overlap=0.25, mode="gaussian",
                                              return_logits=True)
    assert ((logits.cpu() - want).norm() / want.norm()).item() < 1e-4
    assert ((const - want).norm() / want.norm()).item() > 1e-3  # the importance map does change the blend
    assert (labels.cpu().long() != O.squash_predictions(want)).float().mean().item() < 1e-3


def test_graphed_predictor_equals_eager():
    """Forward pass replayed from a CUDA graph (full batches) + eager ragged last batch == eager predictor,
    bit for bit, also after the weights changed (the repack is part of the graph)."""
    from ct_image_segmentation_b200.inference import GraphedPredictor
    torch.manual_seed(2)
    net = B.UNet(3, 1, 10, [16, 32, 64], [2, 2], num_res_units=2, dtype=torch.bfloat16).to(DEV)
    x = torch.randn(1, 1, 40, 40, 24, device=DEV)
    roi = (16, 16, 16)
    assert len(window_list(x.shape[2:], roi, 0.25)) % 4 != 0  # the last batch is ragged
    pred = GraphedPredictor(net, torch.zeros(4, 1, *roi, device=DEV))
    for _ in range(2):
        lab_e, log_e = sliding_window_inference(x, roi, 4, net, 0.25, return_logits=True, rank=0, world=1)
        lab_g, log_g = sliding_window_inference(x, roi, 4, pred, 0.25, return_logits=True, rank=0, world=1)
        assert torch.equal(lab_e, lab_g) and torch.equal(log_e, log_g)
        with torch.no_grad():
            for p in net.parameters():
                p.mul_(1.01)


def test_small_volume_is_padded_to_roi():
    net = B.UNet(3, 1, 10, [8, 16, 16], [2, 2], num_res_units=1, dtype=torch.float32).to(DEV)
    x = torch.randn(1, 1, 12, 20, 16, device=DEV)
    labels = sliding_window_inference(x, (16, 16, 16), 1, net)
    assert tuple(labels.shape) == (1, 12, 20, 16)


def test_patch_sampler_matches_oracle_windowing():
    g = torch.Generator().manual_seed(3)
    hu = torch.randint(-1024, 3072, (40, 56, 48), generator=g, dtype=torch.int16)
    lab = (torch.rand(40, 56, 48, generator=g) > 0.97).to(torch.uint8) * 3
    s = PatchSampler(hu.to(DEV), lab.to(DEV), (32, 32, 32), dtype=torch.float32, rank=1, foreground_prob=0.5)
    img, plab = s.sample(4)
    assert tuple(img.shape) == (4, 1, 32, 32, 32) and tuple(plab.shape) == (4, 32, 32, 32)
    org = s.last_origins.cpu().numpy()
    ref_full = O.window_normalize(hu.numpy(), ("soft_tissue",))[..., 0]
    for b in range(4):
        d, h, w = org[b]
        np.testing.assert_allclose(img[b, 0].cpu().numpy(), ref_full[d:d + 32, h:h + 32, w:w + 32], atol=1e-6, rtol=0)
        assert np.array_equal(plab[b].cpu().numpy(), lab.numpy()[d:d + 32, h:h + 32, w:w + 32])
    # same seed + rank => same patch stream; a different rank draws different patches
    s2 = PatchSampler(hu.to(DEV), lab.to(DEV), (32, 32, 32), dtype=torch.float32, rank=1, foreground_prob=0.5)
    assert np.array_equal(s2.origins(4), org)
    s3 = PatchSampler(hu.to(DEV), lab.to(DEV), (32, 32, 32), dtype=torch.float32, rank=2, foreground_prob=0.5)
    assert not np.array_equal(s3.origins(4), org)


def test_patch_sampler_pads_outside_volume():
    hu = torch.full((20, 24, 24), 100, dtype=torch.int16, device=DEV)
    s = PatchSampler(hu, None, (32, 32, 32), dtype=torch.bfloat16, foreground_prob=0.0)
    img, lab = s.sample(1)
    assert lab is None and tuple(img.shape) == (1, 1, 32, 32, 32)
    inside = O.window_normalize(np.array([100], dtype=np.int16), ("soft_tissue",))[0, 0]
    outside = O.window_normalize(np.array([-1024], dtype=np.int16), ("soft_tissue",))[0, 0]
    o = s.last_origins.cpu().numpy()[0]
    assert (o <= 0).all()
    v = img[0, 0].float().cpu()
    assert abs(v[-o[0], -o[1], -o[2]].item() - inside) < 2e-2
    assert abs(v[31, 31, 31].item() - outside) < 2e-2


def test_patient_npz_to_gpu_sampler(tmp_path):
    """Reference .npz patient file -> resident volume + fused mask squashing -> GPU patch sampler."""
    import numpy as np
    from ct_image_segmentation_b200 import data
    rng = np.random.default_rng(3)
    vol = rng.integers(-1000, 2000, size=(1, 20, 24, 28)).astype(np.int16)
    masks = np.zeros((9, 20, 24, 28), dtype=np.uint8)
    for c in range(9):
        masks[c, 2 * c:2 * c + 3, 4:12, 6:16] = 1
    np.savez(tmp_path / "p.npz", image=vol, masks=masks, mask_indicator=np.ones(9))
    pat = data.load_patient_npz(tmp_path / "p.npz")
    smp = pat.sampler((16, 16, 16), window="soft_tissue", dtype=torch.float32, foreground_prob=1.0)
    img, lab = smp.sample(3)
    assert img.shape == (3, 1, 16, 16, 16) and lab.shape == (3, 16, 16, 16) and lab.dtype == torch.uint8
    full = torch.from_numpy(pat.labels())
    for b, (d0, h0, w0) in enumerate(smp.last_origins.cpu().tolist()):
        ref = full[d0:d0 + 16, h0:h0 + 16, w0:w0 + 16]
        assert torch.equal(lab[b].cpu()[:ref.shape[0], :ref.shape[1], :ref.shape[2]], ref)
    assert int(lab.max()) > 0
