"""Data-parallel training plumbing: one process per GPU, one flat gradient all-reduce per step.

The reference has no in-repo parallelism; multi-GPU training is PyTorch-Lightning DDP
(``Trainer.from_argparse_args`` at reference ``capstone/volumetric/base_trainer.py:196``), i.e.
NCCL all-reduce of gradient buckets averaged over ranks.  InstanceNorm is per sample and the Dice
loss is averaged per rank, so with equal local batch sizes averaging the gradients over ranks
reproduces DDP exactly.  This module does that with ONE flat fp32 bucket
(4.8 M parameters = 19 MB for the 16-256 net) over NCCL / NVLink; ``gloo`` works for CPU tests.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> tuple:
    """(rank, world, local_rank) from the torchrun environment; no-op for a single process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


class GradientBucket:
    """Flat fp32 bucket over a fixed parameter list: gather grads -> all-reduce(sum) -> /world ->
    scatter back into ``p.grad``."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.sizes = [p.numel() for p in self.params]
        self.flat = torch.zeros(sum(self.sizes), dtype=torch.float32, device=dev)
        self.views = [v.view_as(p) for v, p in zip(self.flat.split(self.sizes), self.params)]

    def allreduce_mean(self, group=None) -> None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        torch._foreach_copy_(self.views, grads)
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            self.flat.mul_(1.0 / world)
        for p, v in zip(self.params, self.views):
            p.grad = v  # parameters now read their gradient straight from the bucket


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Round-robin assignment of independent work items (patches, windows) to ranks."""
    return list(range(rank, n_items, world))
