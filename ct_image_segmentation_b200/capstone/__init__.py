"""Drop-in mirror of the reference's ``capstone`` package for the hot path only.

``from ct_image_segmentation_b200.capstone.models import UNet, MultipleLossWrapper,
DiceMetricWrapper`` matches reference ``capstone/models/__init__.py:1-3``; the LightningModule
surface lives in ``.volumetric.base_trainer`` / ``.training.base_trainer``.  Data loading, W&B
callbacks, NRRD utilities and the CLI glue are out of scope.
"""
