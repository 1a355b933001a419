"""Whole-step CUDA-graph execution: forward + loss + backward captured once, replayed per step.

A training step of the 16-256 U-Net is ~250 kernel launches of a few microseconds each; issued
eagerly from Python they are host-bound ("Scaling >= 7x").  All b200seg
launches go to the current stream, allocate nothing themselves and never synchronise, so the
step can be captured into one CUDA graph (activations live in the graph's private pool, TMA
descriptors are encoded at capture time for those fixed addresses).  The packed weights are
refreshed INSIDE the graph (the pack kernels are captured), so optimiser updates between
replays are picked up.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch
import torch.distributed as dist

from .parallel import GradientBucket
from .unet import UNet


class GraphedTrainStep:
    """``step(images, labels) -> loss`` with forward + loss + backward + the gradient all-reduce of the
    data-parallel group replayed from a CUDA graph, then (eagerly) the optimiser.

    The weight-gradient kernels write straight into the flat fp32 bucket (``UNet.bind_grad_sink``): nothing is
    gathered after the backward pass.  With more than one rank the bucket range holding everything but the two
    outermost down layers (99 % of the parameters; complete when the reverse plan returns to the two
    high-resolution encoder levels) is all-reduced on a communication stream BESIDE the rest of the backward
    pass; only that small first range is reduced after it.

    ``images`` / ``labels`` may live on the host (pinned) or on the device; they are copied into
    the graph's static input buffers.  The returned loss is a device scalar (no host sync).
    """

    def __init__(self, net: UNet, loss_fn: Callable, optimizer: Optional[torch.optim.Optimizer],
                 images: torch.Tensor, labels: torch.Tensor, warmup: int = 2, use_graph: bool = True,
                 overlap_allreduce: Optional[bool] = None, metric: Optional[Callable] = None):
        """``metric`` (optional): the reference runs its Dice METRIC on every training step (``_log_dice_scores``,
        ``capstone/volumetric/base_trainer.py:116-132``: clone, softmax, argmax, two one-hots, sums).  Either
        ``"fused"`` -- ``loss_fn`` is a ``DiceLoss(with_metric=True)`` whose forward kernel counts {tp, |pred|,
        |target|} in the SAME pass over the logits (no extra HBM traffic); only the (B, 9) epilogue
        ``dice_from_counts`` is added, on a second stream -- or a callable ``metric(logits, labels) -> tensors``
        (e.g. ``DiceMetricWrapper().from_logits``: one fused argmax + integer-count launch) issued on that second
        stream, a branch BESIDE the backward pass in the captured graph.  The result is in ``self.metric_out``."""
        if metric == "fused" and not getattr(loss_fn, "with_metric", False):
            raise ValueError('metric="fused" needs loss_fn = DiceLoss(..., with_metric=True)')
        self.net, self.loss_fn, self.optimizer = net, loss_fn, optimizer
        self.metric, self.metric_out = metric, None
        self._metric_stream = None
        dev = next(net.parameters()).device
        self.static_images = torch.empty(images.shape, dtype=images.dtype, device=dev)
        self.static_labels = torch.empty(labels.shape, dtype=labels.dtype, device=dev)
        self.static_images.copy_(images)
        self.static_labels.copy_(labels)
        self.bucket = GradientBucket(net.parameters())
        if hasattr(optimizer, "bind_grad_buffer"):  # FlatAdam reads the all-reduced bucket directly
            optimizer.bind_grad_buffer(self.bucket.flat, self.bucket.params)
        self._sink = dict(zip(self.bucket.params, self.bucket.views))
        for p, v in zip(self.bucket.params, self.bucket.views):
            p.grad = v  # the optimiser reads the (all-reduced) bucket
        self._world = dist.get_world_size() if dist.is_initialized() else 1
        if overlap_allreduce is None:
            overlap_allreduce = os.environ.get("B200SEG_OVERLAP_ALLREDUCE", "1") == "1"
        self._deep = self._deep_range(net) if (self._world > 1 and overlap_allreduce) else None
        self._comm_stream = torch.cuda.Stream(device=dev) if self._deep is not None else None
        self._deep_issued = False
        # graph mode: the exchange is part of the captured step (B200SEG_EXCHANGE_IN_GRAPH=0: after the replay)
        self._exchange_in_graph = os.environ.get("B200SEG_EXCHANGE_IN_GRAPH", "1") == "1"
        if use_graph and not self._exchange_in_graph:
            self._deep = None
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.loss: Optional[torch.Tensor] = None
        self.use_graph = use_graph
        self.launches_per_step = self.tc_launches_per_step = None
        self._pad_pool = {}
        # host inputs: double-buffered staging filled on a copy stream, so the H2D transfer of step
        # k+1 overlaps the compute of step k (the static inputs are live for the whole step: the
        # first-layer wgrad reads them at the very end of the backward)
        self._copy_stream = torch.cuda.Stream(device=dev)
        self._stage = [(torch.empty_like(self.static_images), torch.empty_like(self.static_labels))
                       for _ in range(2)]
        self._ready = [torch.cuda.Event() for _ in range(2)]
        self._consumed = [torch.cuda.Event() for _ in range(2)]
        self._k = 0
        if use_graph:
            net.enable_wgrad_stream(True)  # weight gradients as a parallel branch of the captured graph
            self._capture(warmup)

    # ---- gradient exchange -------------------------------------------------------------------------
    def _deep_range(self, net: UNet):
        """(depth, lo, hi): the flat-bucket range [lo, hi) that is all-reduced as soon as the reverse plan
        has issued the weight gradients of level ``depth``.  ``parameters()`` walks the module tree depth
        first -- down layers, then the levels below, then the up layers -- and the reverse plan runs the up
        layers first, so when level ``depth`` is done everything FROM its first parameter TO THE END of the
        bucket is complete: one contiguous range; only the down layers above it (a few 10^4 parameters)
        remain for after the backward pass.  Depth 2 leaves the two full-resolution encoder levels (a third of
        the backward pass) to hide the exchange behind; shallower nets use depth 1."""
        levels, m = 0, net.model
        while isinstance(m, torch.nn.Sequential) and len(m) == 3 and hasattr(m[1], "submodule"):
            levels, m = levels + 1, m[1].submodule
        if levels < 1:
            return None
        depth = 2 if levels >= 3 else 1
        outer, m = [], net.model
        for _ in range(depth):  # down layers of the levels above `depth`: the first parameters of the bucket
            outer += list(m[0].parameters())
            m = m[1].submodule
        n_outer = len(outer)
        if [id(p) for p in self.bucket.params[:n_outer]] != [id(p) for p in outer]:
            return None  # not the depth-first order (e.g. frozen parameters): plain exchange
        lo = sum(self.bucket.sizes[:n_outer])
        return depth, lo, self.bucket.flat.numel()

    def _allreduce_mean(self, t: torch.Tensor) -> None:
        if t.numel() == 0:
            return
        if dist.get_backend() == "nccl":  # the average is taken inside the collective
            dist.all_reduce(t, op=dist.ReduceOp.AVG)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            t.mul_(1.0 / self._world)

    def _on_level_done(self, depth: int) -> None:
        """Reverse-plan hook (runs on the autograd thread, on the step's stream): level ``depth`` and below
        have issued all their launches -> all-reduce their bucket range beside the rest of the backward."""
        if self._deep is None or depth != self._deep[0]:
            return
        comm = self._comm_stream
        comm.wait_stream(torch.cuda.current_stream())
        if self.net.wgrad_stream and self.net._wgrad_side is not None:
            comm.wait_stream(self.net._wgrad_side)  # the weight gradients are written on the second stream
        with torch.cuda.stream(comm):
            self._allreduce_mean(self.bucket.flat[self._deep[1]:self._deep[2]])
        self._deep_issued = True

    def _exchange_rest(self) -> None:
        if self._world == 1:
            return
        if self._deep_issued:
            self._allreduce_mean(self.bucket.flat[:self._deep[1]])
            torch.cuda.current_stream().wait_stream(self._comm_stream)
        else:
            self._allreduce_mean(self.bucket.flat)

    def _loss_and_metric(self):
        logits = self.net(self.static_images)
        if self.metric == "fused":
            from .metrics import dice_from_counts
            loss = self.loss_fn(logits, self.static_labels.unsqueeze(1))
            counts = self.loss_fn.metric_counts
            if self._metric_stream is None:
                self._metric_stream = torch.cuda.Stream(device=logits.device)
            self._metric_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._metric_stream), torch.no_grad():
                self.metric_out = dice_from_counts(counts)
            self._metric_logits = counts
            return loss
        if self.metric is not None:
            # the metric only reads the logits: a side branch, joined at the end of the step
            if self._metric_stream is None:
                self._metric_stream = torch.cuda.Stream(device=logits.device)
            cur = torch.cuda.current_stream()
            self._metric_stream.wait_stream(cur)
            with torch.cuda.stream(self._metric_stream), torch.no_grad():
                self.metric_out = self.metric(logits.detach(), self.static_labels)
            self._metric_logits = logits  # stays referenced until the join (caching allocator)
        return self.loss_fn(logits, self.static_labels.unsqueeze(1))

    def _fwd_bwd(self, exchange: bool = True):
        from . import ops
        self._deep_issued = False
        self.net.bind_grad_sink(self._sink, self._on_level_done if (self._deep is not None and exchange) else None)
        try:
            if self.use_graph:
                # zero-padded 10-class buffers are allocated (and zeroed) once and reused by every replay:
                # a replayed step's activations are dead before the next replay starts
                with ops.padded_buffer_pool(self._pad_pool):
                    loss = self._loss_and_metric()
                    loss.backward()
            else:
                loss = self._loss_and_metric()
                loss.backward()
        finally:
            self.net.bind_grad_sink(None)
        if self.metric is not None:
            torch.cuda.current_stream().wait_stream(self._metric_stream)
            self._metric_logits = None
        if exchange and not (self.use_graph and not self._exchange_in_graph and torch.cuda.is_current_stream_capturing()):
            self._exchange_rest()
        return loss

    def local_gradients(self) -> torch.Tensor:
        """This rank's OWN gradient (no exchange) for the current inputs and parameters, as a copy of the flat bucket:
        one eager forward + backward.  With ``check_exchange`` it is the evidence that the in-graph, overlapped
        exchange leaves the mean of the ranks' local gradients in the bucket."""
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # (not on the default stream: keeps autograd's stream bookkeeping off it)
            self._fwd_bwd(exchange=False)
            out = self.bucket.flat.clone()
        torch.cuda.current_stream().wait_stream(side)
        return out

    def check_exchange(self) -> dict:
        """Data-parallel proof, outside any timed region: (1) local gradients without exchange, averaged over the
        ranks with ONE plain all-reduce; (2) the step as it is timed (graph replay incl. the overlapped in-graph
        exchange), no optimiser update in between.  Returns the largest deviation between the two buckets."""
        opt, self.optimizer = self.optimizer, None
        try:
            want = self.local_gradients()
            if self._world > 1:
                self._allreduce_mean(want)
            self.__call__(None, None)
            torch.cuda.synchronize()
            got = self.bucket.flat
            diff = (got - want).abs().max()
            scale = want.abs().max().clamp_min(1e-30)
            return {"world": self._world, "max_abs_diff": float(diff), "max_abs_grad": float(scale),
                    "max_rel_to_largest": float(diff / scale), "bitwise_equal": bool(torch.equal(got, want)),
                    "n_params": int(got.numel())}
        finally:
            self.optimizer = opt

    def _capture(self, warmup: int):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the default stream, as graph capture requires
            for _ in range(max(1, warmup)):  # (also brings the NCCL communicator up before the capture)
                self._fwd_bwd()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.net.reset_packed_cache()  # the weight-repack kernels must be part of the graph
        from . import _lib
        lib = _lib.load()
        before = (lib.b200seg_launch_count(), lib.b200seg_tc_launch_count())
        # capture on a high-priority stream: the serial chain (main branch) is scheduled ahead of the
        # weight-gradient branch, which was created at default (lower) priority
        cap_stream = torch.cuda.Stream(priority=-1) if os.environ.get("B200SEG_GRAPH_PRIO", "1") == "1" else None
        try:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self.loss = self._fwd_bwd()
        except Exception as e:  # noqa: BLE001
            if self._world == 1 or not self._exchange_in_graph:
                raise
            # a communicator that cannot be captured (NCCL build / transport): keep the compute in the graph
            # and run the gradient exchange eagerly after every replay
            import warnings
            warnings.warn(f"b200seg: NCCL all-reduce could not be captured into the CUDA graph ({e}); "
                          f"the gradient exchange runs after the graph replay instead")
            self._exchange_in_graph = False
            self._deep = None
            torch.cuda.synchronize()
            self.net.reset_packed_cache()
            before = (lib.b200seg_launch_count(), lib.b200seg_tc_launch_count())
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=cap_stream):
                self.loss = self._fwd_bwd()
        # b200seg kernels recorded into the graph = launched again on every replay
        self.launches_per_step = lib.b200seg_launch_count() - before[0]
        self.tc_launches_per_step = lib.b200seg_tc_launch_count() - before[1]

    def __call__(self, images: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None):
        if images is not None and not images.is_cuda:
            i = self._k & 1
            self._k += 1
            main = torch.cuda.current_stream()
            with torch.cuda.stream(self._copy_stream):
                self._copy_stream.wait_event(self._consumed[i])
                self._stage[i][0].copy_(images, non_blocking=True)
                if labels is not None:
                    self._stage[i][1].copy_(labels, non_blocking=True)
                self._ready[i].record(self._copy_stream)
            main.wait_event(self._ready[i])
            self.static_images.copy_(self._stage[i][0], non_blocking=True)
            if labels is not None:
                self.static_labels.copy_(self._stage[i][1], non_blocking=True)
            self._consumed[i].record(main)
        else:
            if images is not None:
                self.static_images.copy_(images, non_blocking=True)
            if labels is not None:
                self.static_labels.copy_(labels, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
            loss = self.loss
            if not self._exchange_in_graph:
                self._exchange_rest()
        else:
            loss = self._fwd_bwd()
        if self.optimizer is not None:
            if not hasattr(self.optimizer, "bind_grad_buffer"):
                # a stock torch optimiser reads p.grad: keep it pointing at the bucket even after a
                # zero_grad(set_to_none=True) by the caller
                for p, v in zip(self.bucket.params, self.bucket.views):
                    if p.grad is not v:
                        p.grad = v
            self.optimizer.step()
        return loss
