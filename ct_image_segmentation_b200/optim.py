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
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, amsgrad: bool = False):
        if weight_decay != 0.0 or amsgrad:
            raise NotImplementedError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False), "
                                      "the reference's configuration")
        plist = list(params)
        if plist and isinstance(plist[0], dict):
            raise NotImplementedError("FlatAdam takes ONE parameter group (a flat iterable of parameters)")
        plist = [p for p in plist if p.requires_grad]
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

    def add_param_group(self, param_group):
        if getattr(self, "params", None) is not None:  # (the base constructor adds the one group itself)
            raise NotImplementedError("FlatAdam takes ONE parameter group: the flat buffers are laid out at "
                                      "construction")
        super().add_param_group(param_group)

    def bind_grad_buffer(self, flat_grad: torch.Tensor, params) -> None:
        """Use ``flat_grad`` (same parameter order, e.g. ``GradientBucket.flat``) as the gradient."""
        if [id(p) for p in params] != [id(p) for p in self.params] or flat_grad.numel() != self.flat.numel():
            raise ValueError("gradient buffer does not match the optimiser's parameter list")
        self.flat_grad = flat_grad

    # ---- checkpointing: the moments live in flat private buffers, not in ``self.state`` -----------------
    def state_dict(self):
        """``torch.optim.Adam``'s layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so a checkpoint
        resumes under either optimiser."""
        sd = super().state_dict()
        ids = sd["param_groups"][0]["params"]
        state = {}
        for i, m, v, p in zip(ids, self.exp_avg.split(self.sizes), self.exp_avg_sq.split(self.sizes), self.params):
            state[i] = {"step": torch.tensor(float(self.steps)), "exp_avg": m.view_as(p).clone(),
                        "exp_avg_sq": v.view_as(p).clone()}
        sd["state"] = state
        return sd

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("FlatAdam.load_state_dict: expected one group over the same parameter list")
        if groups[0].get("weight_decay", 0) != 0 or groups[0].get("amsgrad", False):
            raise NotImplementedError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False)")
        for k in ("lr", "betas", "eps"):
            if k in groups[0]:
                self.param_groups[0][k] = groups[0][k]
        state = state_dict.get("state", {})
        steps = set()
        with torch.no_grad():
            for i, m, v, p in zip(groups[0]["params"], self.exp_avg.split(self.sizes),
                                  self.exp_avg_sq.split(self.sizes), self.params):
                st = state.get(i)
                if st is None:  # parameter never stepped
                    m.zero_()
                    v.zero_()
                    continue
                m.view_as(p).copy_(st["exp_avg"])
                v.view_as(p).copy_(st["exp_avg_sq"])
                steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError(f"FlatAdam.load_state_dict: parameters at different steps {sorted(steps)}")
        self.steps = steps.pop() if steps else 0

    @torch.no_grad()
    def step(self, closure=None):
        if closure is not None:
            raise NotImplementedError("FlatAdam does not take a closure")
        g = self.flat_grad
        if g is None:
            g = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self.params])
        grp = self.param_groups[0]
        if grp.get("weight_decay", 0) != 0 or grp.get("amsgrad", False):
            raise NotImplementedError("FlatAdam implements plain Adam (weight_decay=0, amsgrad=False)")
        self.steps += 1
        lib = _lib.load()
        _lib.check(lib.b200seg_adam_step(self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                         self.exp_avg_sq.data_ptr(), self.flat.numel(), float(grp["lr"]),
                                         float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                                         self.steps, torch.cuda.current_stream().cuda_stream), "b200seg_adam_step")
        # the kernel wrote the parameters through a raw pointer: tell autograd / the packed-weight caches
        # (UNet._pack / _repack_all key on ``_version``) that every parameter changed
        for p in self.params:
            torch.autograd.graph.increment_version(p)
        return None
