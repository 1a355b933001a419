"""Sliding-window inference with the windows sharded over the GPUs of one box.

Absent from the reference (it nearest-resizes whole volumes to 96x256x256,
``capstone/volumetric/transforms.py:9-23``); semantics are MONAI's
``sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25,
mode="constant")``: symmetric zero padding up to the ROI, scan interval
``int(roi * (1 - overlap))``, last window shifted back to the border, constant importance map,
``out = sum(window logits) / count``.  Windows are independent forward passes: rank r takes
windows r, r+world, ...; the fp32 accumulators are summed with ONE all-reduce, then every rank
averages + arg-maxes (fused kernel) -- no halo exchange.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops
from .parallel import shard_indices


class GraphedPredictor:
    """``predictor`` for :func:`sliding_window_inference`: the network's forward pass for ONE batch shape
    (``sw_batch_size`` windows of the ROI) captured into a CUDA graph and replayed per batch -- a forward pass
    of the 16-256 U-Net is ~85 launches of a few microseconds, host-bound when issued from Python.  Batches
    of another shape (the last, ragged one) run eagerly.  The returned tensor is the graph's output buffer:
    consume it (as ``sliding_window_inference`` does) before the next call."""

    def __init__(self, net, example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("b200seg inference runs on CUDA tensors only")
        self.net = net
        self.static_in = example.clone()
        self._pool: dict = {}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():  # warm-up off the default stream, as capture requires
            for _ in range(max(1, warmup)):
                with ops.padded_buffer_pool(self._pool):
                    net(self.static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        if hasattr(net, "reset_packed_cache"):
            net.reset_packed_cache()  # the weight-repack launch must be part of the graph
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad(), ops.padded_buffer_pool(self._pool):
            self.static_out = net(self.static_in)

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        if batch.shape != self.static_in.shape or batch.dtype != self.static_in.dtype:
            with torch.no_grad():
                return self.net(batch)
        self.static_in.copy_(batch)
        self.graph.replay()
        return self.static_out


def scan_starts(size: int, roi: int, overlap: float) -> List[int]:
    """Window start offsets along one axis."""
    if roi >= size:
        return [0]
    interval = max(int(roi * (1 - overlap)), 1)
    n = 1
    while (n - 1) * interval + roi < size:
        n += 1
    return [min(k * interval, size - roi) for k in range(n)]


def window_list(dims: Sequence[int], roi: Sequence[int], overlap: float) -> List[Tuple[int, int, int]]:
    st = [scan_starts(s, r, overlap) for s, r in zip(dims, roi)]
    return [(a, b, c) for a in st[0] for b in st[1] for c in st[2]]


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.25,
                             return_logits: bool = False, rank: Optional[int] = None,
                             world: Optional[int] = None, partial_only: bool = False):
    """``inputs`` (1, Cin, D, H, W) on the GPU -> uint8 label map (1, D, H, W) (and, with
    ``return_logits``, the averaged fp32 logits (1, C, D, H, W))."""
    if inputs.dim() != 5 or inputs.shape[0] != 1:
        raise ValueError("sliding_window_inference takes one 3-D volume: (1, C, D, H, W)")
    if not inputs.is_cuda:
        raise RuntimeError("b200seg inference runs on CUDA tensors only")
    lib = _lib.load()
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
        rank = dist.get_rank() if dist.is_initialized() else 0
    dims = tuple(inputs.shape[2:])
    pad = []
    for sz, r in zip(reversed(dims), reversed(tuple(roi_size))):
        diff = max(r - sz, 0)
        pad += [diff // 2, diff - diff // 2]
    x = F.pad(inputs, pad) if any(pad) else inputs
    pd, ph, pw = x.shape[2:]
    wins = window_list((pd, ph, pw), roi_size, overlap)
    mine = [wins[i] for i in shard_indices(len(wins), rank, world)]
    acc = cnt = None
    n_classes = None
    stream = torch.cuda.current_stream().cuda_stream
    for i in range(0, len(mine), sw_batch_size):
        chunk = mine[i:i + sw_batch_size]
        batch = torch.cat([x[:, :, a:a + roi_size[0], b:b + roi_size[1], c:c + roi_size[2]] for a, b, c in chunk], 0)
        with torch.no_grad():
            pred = predictor(batch)
        cl = ops.to_channels_last(pred)  # (B, d, h, w, C), a view for UNet outputs
        _, wd, wh, ww, c, ld = ops.cl_info(cl)
        if acc is None:
            n_classes = c
            acc = torch.zeros(pd, ph, pw, c, dtype=torch.float32, device=x.device)
            cnt = torch.zeros(pd, ph, pw, dtype=torch.float32, device=x.device)
        for j, (a, b, c0) in enumerate(chunk):
            _lib.check(lib.b200seg_window_accumulate(ops.dtype_code(cl.dtype), cl[j].data_ptr(), ld, acc.data_ptr(),
                                                     cnt.data_ptr(), n_classes, wd, wh, ww, pd, ph, pw, a, b, c0,
                                                     stream), "b200seg_window_accumulate")
    if acc is None:  # this rank got no window (more ranks than windows): contribute zeros
        probe = predictor(x[:, :, :roi_size[0], :roi_size[1], :roi_size[2]])
        n_classes = probe.shape[1]
        acc = torch.zeros(pd, ph, pw, n_classes, dtype=torch.float32, device=x.device)
        cnt = torch.zeros(pd, ph, pw, dtype=torch.float32, device=x.device)
    if partial_only:  # caller combines the shards itself (tests; custom reductions)
        return acc, cnt
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    return _finalize(acc, cnt, dims, roi_size, return_logits)


def _finalize(acc, cnt, dims, roi_size, return_logits):
    """Average + arg-max of the (summed) accumulators, cropped back to the unpadded volume."""
    lib = _lib.load()
    pd, ph, pw, n_classes = acc.shape
    stream = torch.cuda.current_stream().cuda_stream
    x = acc
    labels = torch.empty(pd, ph, pw, dtype=torch.uint8, device=x.device)
    mean = torch.empty_like(acc) if return_logits else None
    _lib.check(lib.b200seg_accum_argmax(acc.data_ptr(), cnt.data_ptr(), labels.data_ptr(),
                                        None if mean is None else mean.data_ptr(), pd * ph * pw, n_classes, stream),
               "b200seg_accum_argmax")
    sl = tuple(slice(max(r - s, 0) // 2, max(r - s, 0) // 2 + s) for s, r in zip(dims, roi_size))
    labels = labels[sl].unsqueeze(0)
    if return_logits:
        return labels, mean[sl].permute(3, 0, 1, 2).unsqueeze(0)
    return labels
