"""``BaseUNet3D`` with the reference's constructor / step contract on the B200 hot path.

Mirrors reference ``capstone/volumetric/base_trainer.py:21-182``: same constructor arguments and
hyper-parameter names, ``forward``, ``training_step``, ``validation_step``, ``_shared_step`` ->
``(images, masks, mask_indicator, prediction, total_loss)``, ``configure_optimizers`` -> Adam,
``add_model_specific_args``, verbatim metric names.  Differences, all on purpose:
 * the loss dispatcher implements the INTENDED 3-D behaviour (the shipped
   ``MultipleLossWrapper3D`` resolves the 2-D table and asserts on 3-D targets);
 * ``_log_dice_scores`` uses one fused argmax+count pass over the logits instead of
   ``clone -> softmax -> argmax -> 2x one_hot -> sums``;
 * extra keyword ``dtype`` (``torch.bfloat16`` production / ``torch.float32`` check mode).
"""
from argparse import ArgumentParser
from typing import List

import torch
import torch.optim as optim

from ...losses import N_CLASSES, STRUCTURES, MultipleLossWrapper3D
from ...metrics import DiceMetricWrapper3D, squash_masks
from ...unet import UNet
from .._lightning import LightningModule

SEED = 12342


class BaseUNet3D(LightningModule):
    def __init__(self, filters: List = [16, 32, 64, 128, 256], use_res_units: bool = False,
                 downsample: bool = False, lr: float = 1e-3, loss_fx: list = ["CrossEntropy"],
                 exclude_missing: bool = False, dtype: torch.dtype = torch.bfloat16, **kwargs) -> None:
        super().__init__()
        assert isinstance(loss_fx, list), "This module expects a list of loss functions"
        loss_fx.sort()  # consistent order of loss functions (reference :35)
        kwargs.setdefault("batch_size", 1)
        kwargs.setdefault("transform_degree", 0)
        self.save_hyperparameters("batch_size", "transform_degree", "filters", "use_res_units",
                                  "downsample", "lr", "loss_fx", "exclude_missing")
        self._compute_dtype = dtype
        self.unet = self._construct_model()
        self.loss_func = MultipleLossWrapper3D(losses=loss_fx, exclude_missing=exclude_missing)
        self.dice_score = DiceMetricWrapper3D()
        # the per-step Dice metric (reference :116-132) rides on the Dice loss's own pass over the logits when Dice
        # is among the losses and takes the shared multi-loss pass otherwise (then: one fused argmax + count launch)
        self._fused_metric = ("Dice" in loss_fx and len(loss_fx) == 1 and not exclude_missing
                              and self.loss_func.enable_metric_counts())

    @property
    def _n_classes(self):
        return len(STRUCTURES) + 1  # additional background

    def _construct_model(self):
        # reference :62-72 -- strides, in_channels and num_res_units are hard-wired
        return UNet(dimensions=3, in_channels=1, out_channels=self._n_classes,
                    channels=self.hparams.filters, strides=[2, 2, 2, 2], num_res_units=2,
                    dtype=self._compute_dtype)

    def forward(self, x):
        return self.unet(x)

    def training_step(self, batch, batch_idx):
        _, _, _, _, loss = self._shared_step(batch, is_training=True)
        return loss

    def validation_step(self, batch, batch_idx):
        self._shared_step(batch, is_training=False)

    def _shared_step(self, batch, is_training: bool):
        (images, masks, mask_indicator) = batch
        masks = squash_masks(masks, self._n_classes)  # (B, *S) labels 0..9
        mask_indicator = mask_indicator.type_as(images)
        prefix = "train" if is_training else "val"
        prediction = self.forward(images)
        loss_dict = self.loss_func(input=prediction, target=masks, mask_indicator=mask_indicator)
        total_loss = torch.stack(list(loss_dict.values())).sum()
        for name, loss_value in loss_dict.items():
            self.log(f"{name} Loss ({prefix})", loss_value, on_step=False, on_epoch=True)
        self._log_dice_scores(prediction, masks, mask_indicator, prefix)
        return images, masks, mask_indicator, prediction, total_loss

    def configure_optimizers(self):
        return optim.Adam(self.parameters(), lr=self.hparams.lr)

    def _log_dice_scores(self, prediction, masks, mask_indicator, prefix):
        self.eval()
        with torch.no_grad():
            counts = self.loss_func.metric_counts if self._fused_metric else None
            if counts is not None:
                from ...metrics import dice_from_counts
                dice_mean, dice_per_class = dice_from_counts(counts)
            else:
                dice_mean, dice_per_class = self.dice_score.from_logits(prediction.detach(), masks)
            for structure, score in zip(STRUCTURES, dice_per_class):
                self.log(f"{structure} Dice ({prefix})", score, on_step=False, on_epoch=True)
            self.log(f"Mean Dice Score ({prefix})", dice_mean, on_step=False, on_epoch=True)
        self.train()

    @staticmethod
    def add_model_specific_args(parent_parser):
        parser = ArgumentParser(parents=[parent_parser], add_help=False)
        parser.add_argument("--batch_size", type=int, default=1, help="Batch size")
        parser.add_argument("--transform_degree", type=int, default=0)
        parser.add_argument("--filters", nargs=5, type=int, default=[64, 128, 256, 512, 1024])
        parser.add_argument("--use_res_units", action="store_true", default=False)
        parser.add_argument("--downsample", action="store_true", default=False)
        parser.add_argument("--lr", type=float, default=1e-3, help="Learning rate")
        parser.add_argument("--loss_fx", nargs="+", type=str, default="CrossEntropy")
        parser.add_argument("--exclude_missing", action="store_true", default=False)
        return parser
