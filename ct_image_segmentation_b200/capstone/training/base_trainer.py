"""``BaseUNet2D`` with the reference's constructor / step contract on the B200 hot path
(reference ``capstone/training/base_trainer.py:22-209``).  The optional 3->1 ``conv1x1`` stays a
``torch.nn.Conv2d`` (two multiply-adds per pixel in front of the network; not on the hot path)."""
from argparse import ArgumentParser
from typing import List

import torch
import torch.nn as nn
import torch.optim as optim

from ...losses import STRUCTURES, MultipleLossWrapper
from ...metrics import DiceMetricWrapper, squash_masks
from ...unet import UNet
from .._lightning import LightningModule

SEED = 12342


class BaseUNet2D(LightningModule):
    def __init__(self, filters: List = [64, 128, 256, 512, 1024], use_res_units: bool = False,
                 downsample: bool = False, lr: float = 1e-3, loss_fx: list = ["Focal", "Dice"],
                 exclude_missing: bool = False, dtype: torch.dtype = torch.bfloat16, **kwargs) -> None:
        super().__init__()
        assert isinstance(filters, list)
        assert len(filters) == 5, "This module requires a standard 5 block UNet specification"
        assert isinstance(loss_fx, list), "This module expects a list of loss functions"
        loss_fx.sort()
        kwargs.setdefault("batch_size", 1)
        kwargs.setdefault("transform_degree", 0)
        self.save_hyperparameters("batch_size", "transform_degree", "filters", "use_res_units",
                                  "downsample", "lr", "loss_fx", "exclude_missing")
        self._compute_dtype = dtype
        self.conv1x1 = nn.Conv2d(in_channels=3, out_channels=1, kernel_size=1, stride=1)
        self.unet = self._construct_model()
        self.loss_func = MultipleLossWrapper(losses=loss_fx, exclude_missing=exclude_missing)
        self.dice_score = DiceMetricWrapper()

    @property
    def _n_classes(self):
        return len(STRUCTURES) + 1

    def _construct_model(self):
        in_channels = 1 if (self.hparams.downsample or (self.hparams.transform_degree == 0)) else 3
        return UNet(dimensions=2, in_channels=in_channels, out_channels=self._n_classes,
                    channels=self.hparams.filters, strides=[2, 2, 2, 2],
                    num_res_units=(2 if self.hparams.use_res_units else 0), dtype=self._compute_dtype)

    def forward(self, x):
        if self.hparams.downsample:
            x = self.conv1x1(x)
        return self.unet(x)

    def training_step(self, batch, batch_idx):
        _, _, _, _, loss = self._shared_step(batch, prefix="train")
        return loss

    def validation_step(self, batch, batch_idx):
        self._shared_step(batch, prefix="val")

    def test_step(self, batch, batch_idx):
        self._shared_step(batch, prefix="test")

    def _shared_step(self, batch, prefix: str):
        images, masks, mask_indicator, *dist_maps = batch
        masks = squash_masks(masks, self._n_classes)
        mask_indicator = mask_indicator.type_as(images)
        prediction = self.forward(images)
        dist_maps = None if (len(dist_maps) == 0) else dist_maps[0]  # reference :101 (Boundary loss input)
        loss_dict = self.loss_func(input=prediction, target=masks, mask_indicator=mask_indicator,
                                   dist_maps=dist_maps)
        total_loss = torch.stack(list(loss_dict.values())).sum()
        for name, loss_value in loss_dict.items():
            self.log(f"{name} Loss ({prefix})", loss_value, on_step=False, on_epoch=True)
        self._log_dice_scores(prediction, masks, mask_indicator, prefix)
        return images, masks, mask_indicator, prediction, total_loss

    def _log_dice_scores(self, prediction, masks, mask_indicator, prefix):
        self.eval()
        with torch.no_grad():
            pred = prediction.detach()
            if self.hparams.exclude_missing:  # reference :123-125, no indicator for background
                pred = pred.clone()
                pred[:, 1:] = pred[:, 1:] * mask_indicator[:, :, None, None].to(pred.dtype)
            dice_mean, dice_per_class = self.dice_score.from_logits(pred, masks)
            for structure, score in zip(STRUCTURES, dice_per_class):
                self.log(f"{structure} Dice ({prefix})", score, on_step=False, on_epoch=True)
            self.log(f"Mean Dice Score ({prefix})", dice_mean, on_step=False, on_epoch=True)
        self.train()

    def configure_optimizers(self):
        optimizer = optim.Adam(self.parameters(), lr=self.hparams.lr)
        scheduler = optim.lr_scheduler.ReduceLROnPlateau(optimizer, mode="max", factor=0.5, threshold=0.01)
        return {"optimizer": optimizer, "lr_scheduler": scheduler, "monitor": "Mean Dice Score (val)"}

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = ArgumentParser(parents=[parent_parser], add_help=False)
        parser.add_argument("--batch_size", type=int, default=128)  # reference capstone/training/base_trainer.py:155
        parser.add_argument("--transform_degree", type=int, default=0)
        parser.add_argument("--filters", nargs=5, type=int, default=[64, 128, 256, 512, 1024])
        parser.add_argument("--use_res_units", action="store_true", default=False)
        parser.add_argument("--downsample", action="store_true", default=False)
        parser.add_argument("--lr", type=float, default=1e-3)
        parser.add_argument("--loss_fx", nargs="+", type=str, default=["Focal", "Dice"])
        parser.add_argument("--exclude_missing", action="store_true", default=False)
        return parser
