"""``pl.LightningModule`` when pytorch_lightning is importable, else a minimal stand-in with the
few members the reference's modules use (``save_hyperparameters``, ``hparams``, ``log``,
``device``), so the hot path can be driven without the training framework."""
from __future__ import annotations

import inspect
from argparse import Namespace

import torch
import torch.nn as nn

try:  # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    LightningModule = pl.LightningModule
    HAVE_LIGHTNING = True
except Exception:  # noqa: BLE001
    HAVE_LIGHTNING = False

    class LightningModule(nn.Module):
        def __init__(self):
            super().__init__()
            self.hparams = Namespace()
            self.logged = {}

        def save_hyperparameters(self, *names):
            frame = inspect.currentframe().f_back
            local = frame.f_locals
            kwargs = local.get("kwargs", {})
            for n in names:
                if n in local:
                    setattr(self.hparams, n, local[n])
                elif n in kwargs:
                    setattr(self.hparams, n, kwargs[n])
                else:
                    raise KeyError(f"hyper-parameter {n!r} not passed to {type(self).__name__}")

        def log(self, name, value, **kwargs):
            self.logged[name] = value.detach() if torch.is_tensor(value) else value

        @property
        def device(self):
            try:
                return next(self.parameters()).device
            except StopIteration:
                return torch.device("cpu")
