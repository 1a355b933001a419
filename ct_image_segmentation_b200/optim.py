"""Flat Adam: the reference's optimiser (``torch.optim.Adam(self.parameters(), lr)``,
``capstone/volumetric/base_trainer.py:178-182``) as ONE kernel over a flat fp32 parameter buffer.

The parameters are re-homed into one contiguous buffer (``p.data`` becomes a view, values kept), the
moments are flat, and the gradient is read straight from the flat all-reduce bucket of
``GraphedTrainStep`` -- so a step is one launch moving 7 x 4 bytes per parameter instead of the
multi-tensor launches plus the gather of 63 gradient tensors.
"""
from __future__ import annotations

from typing import Iterable, Optional

import torch

from . import _lib


class FlatAdam(torch.optim.Adam):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8):
        plist = [p for p in params if p.requires_grad]
        if not plist:
            raise ValueError("no trainable parameters")
        if any((not p.is_cuda) or p.dtype != torch.float32 for p in plist):
            raise ValueError("FlatAdam needs CUDA float32 parameters (there is no CPU path)")
        super().__init__(plist, lr=lr, betas=betas, eps=eps)
        self.params = plist
        self.sizes = [p.numel() for p in plist]
        dev = plist[0].device
        self.flat = torch.empty(sum(self.sizes), dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, v in zip(plist, self.flat.split(self.sizes)):
                v = v.view_as(p)
                v.copy_(p.data)
                p.data = v
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.steps = 0
        self.flat_grad: Optional[torch.Tensor] = None

    def bind_grad_buffer(self, flat_grad: torch.Tensor, params) -> None:
        """Use ``flat_grad`` (same parameter order, e.g. ``GradientBucket.flat``) as the gradient."""
        if [id(p) for p in params] != [id(p) for p in self.params] or flat_grad.numel() != self.flat.numel():
            raise ValueError("gradient buffer does not match the optimiser's parameter list")
        self.flat_grad = flat_grad

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FlatAdam does not take a closure")
        g = self.flat_grad
        if g is None:
            g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])
        grp = self.param_groups[0]
        self.steps += 1
        lib = _lib.load()
        _lib.check(lib.b200seg_adam_step(self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), self.flat.numel(), float(grp["lr"]),
                                         float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                                         self.steps, torch.cuda.current_stream().cuda_stream), "b200seg_adam_step")
        return None
