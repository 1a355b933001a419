"""On-disk formats: the reference's .npz patient files and Lightning checkpoints (CPU)."""
import numpy as np
import pytest
import torch

from ct_image_segmentation_b200 import UNet, data
from oracle import monai_ref as O


def test_patient_npz_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    vol = rng.integers(-1024, 3071, size=(1, 6, 10, 12)).astype(np.float32)
    masks = (rng.random((9, 6, 10, 12)) > 0.9).astype(np.uint8)
    ind = np.ones(9)
    ind[3] = 0
    masks[3] = 0
    # exactly what capstone/data/process_miccai.py:_patient_to_3d writes
    np.savez(tmp_path / "0522c0001.npz", image=vol, masks=masks, mask_indicator=ind)
    p = data.load_patient_npz(tmp_path / "0522c0001.npz")
    assert p.patient_id == "0522c0001" and p.image.dtype == np.int16 and p.image.shape == (6, 10, 12)
    assert np.array_equal(p.image, vol[0].astype(np.int16)) and np.array_equal(p.masks, masks)
    assert np.array_equal(p.mask_indicator, ind.astype(np.float32))
    ref = O.squash_masks(torch.from_numpy(masks[None].astype(np.int64)))[0].numpy()
    assert np.array_equal(p.labels(), ref.astype(np.uint8))
    data.save_patient_npz(tmp_path / "again.npz", p.image, p.masks, p.mask_indicator)
    q = data.load_patient_npz(tmp_path / "again.npz")
    assert np.array_equal(q.image, p.image) and np.array_equal(q.masks, p.masks)
    # 2-D slice layout of _patient_to_2d
    # (capstone/data/process_miccai.py:64: slide = vol[:, index] keeps the channel axis: (1, H, W))
    np.savez(tmp_path / "s.npz", image=vol[:, 2], masks=masks[:, 2], mask_indicator=ind)
    s = data.load_patient_npz(tmp_path / "s.npz")
    assert s.image.shape == (1, 10, 12) and s.masks.shape == (9, 1, 10, 12)
    assert np.array_equal(s.image[0], vol[0, 2].astype(np.int16)) and np.array_equal(s.masks[:, 0], masks[:, 2])
    np.savez(tmp_path / "s2.npz", image=vol[0, 2], masks=masks[:, 2], mask_indicator=ind)  # bare (H, W) also accepted
    assert np.array_equal(data.load_patient_npz(tmp_path / "s2.npz").image, s.image)
    np.savez(tmp_path / "s3.npz", image=vol[0, :2], masks=masks[:, 2], mask_indicator=ind)  # 2 "channels": refuse
    with pytest.raises(ValueError):
        data.load_patient_npz(tmp_path / "s3.npz")
    np.savez(tmp_path / "bad.npz", image=vol)
    with pytest.raises(KeyError):
        data.load_patient_npz(tmp_path / "bad.npz")


def test_lightning_checkpoint_import(tmp_path):
    torch.manual_seed(0)
    ref = O.UNet(3, 1, 10, [8, 16, 16], [2, 2], num_res_units=2)
    ckpt = {"state_dict": {"unet." + k: v for k, v in ref.state_dict().items()},
            "hyper_parameters": {"filters": [8, 16, 16], "use_res_units": True, "lr": 1e-3}}
    torch.save(ckpt, tmp_path / "model_large.ckpt")
    net = UNet(3, 1, 10, [8, 16, 16], [2, 2], num_res_units=2, dtype=torch.float32)
    hp = data.load_lightning_checkpoint(tmp_path / "model_large.ckpt", net)
    assert hp["filters"] == [8, 16, 16]
    for (k, a), (k2, b) in zip(net.state_dict().items(), ref.state_dict().items()):
        assert k == k2 and torch.equal(a, b)
    with pytest.raises(KeyError):
        data.unet_state_dict_from_checkpoint({"state_dict": {"other.weight": torch.zeros(1)}})
