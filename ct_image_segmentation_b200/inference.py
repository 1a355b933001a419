"""Sliding-window inference with the windows sharded over the GPUs of one box.

Absent from the reference (it nearest-resizes whole volumes to 96x256x256,
``capstone/volumetric/transforms.py:9-23``); semantics are MONAI's
``sliding_window_inference(inputs, roi_size, sw_batch_size, predictor, overlap=0.25,
mode="constant" | "gaussian")``: symmetric zero padding up to the ROI, scan interval
``int(roi * (1 - overlap))``, last window shifted back to the border, constant or Gaussian importance map,
``out = sum(importance * window logits) / sum(importance)``.

Sharding.  Windows are independent forward passes; the only exchange step is where
windows overlap.  With ``world`` ranks

* the window list (d-major order) is cut into ``world`` contiguous runs: rank r runs the forward pass of run r;
* the OUTPUT volume is cut into ``world`` slabs along its first spatial axis: rank r owns slab r, i.e. it holds
  the only accumulator for those voxels (``slab_d x H x W x C`` fp32 -- 1/world of the volume), averages and
  arg-maxes them (fused kernel) and contributes its uint8 label slab to ONE all-gather (42 MB for 512x512x160);
* a window whose d-range crosses slab borders is cut into runs of d-slices -- contiguous memory in the
  channels-last prediction -- and every run is sent, in the network's own output dtype (bf16: half the bytes of an
  fp32 partial sum), to the slab's owner with point-to-point sends over NVLink (``batch_isend_irecv``);
* the owner adds all runs of its slab **in global window order**, so every voxel sees exactly the additions of the
  single-GPU run in the same order: the label map is bit-identical for every world size.

No whole-volume accumulator is ever all-reduced (round 1 did: 1.7 GB of fp32 for this volume).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F

from . import _lib, ops


# ---- predictor ----------------------------------------------------------------------------------------
class GraphedPredictor:
    """``predictor`` for :func:`sliding_window_inference`: the network's forward pass captured into CUDA graphs,
    one per batch size (``sw_batch_size`` windows of the ROI, and lazily the smaller ragged batches), replayed per
    batch -- a forward pass of the 16-256 U-Net is ~85 launches of a few microseconds, host-bound when issued from
    Python.  Batches of another spatial shape run eagerly.  The returned tensor is the graph's output buffer:
    consume it (as ``sliding_window_inference`` does) before the next call."""

    def __init__(self, net, example: torch.Tensor, warmup: int = 2):
        if not example.is_cuda:
            raise RuntimeError("b200seg inference runs on CUDA tensors only")
        self.net = net
        self.out_channels = getattr(net, "out_channels", None)
        self.compute_dtype = getattr(net, "compute_dtype", None)
        self._warmup = warmup
        self._graphs: Dict[int, tuple] = {}
        self._shape = tuple(example.shape[1:])
        self._dtype = example.dtype
        self._capture(example.clone())

    def _capture(self, static_in: torch.Tensor):
        net, pool = self.net, {}
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():  # warm-up off the default stream, as capture requires
            for _ in range(max(1, self._warmup)):
                with ops.padded_buffer_pool(pool):
                    net(static_in)
        cur.wait_stream(side)
        torch.cuda.synchronize()
        if hasattr(net, "reset_packed_cache"):
            net.reset_packed_cache()  # the weight-repack launch must be part of the graph
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad(), ops.padded_buffer_pool(pool):
            static_out = net(static_in)
        self._graphs[static_in.shape[0]] = (graph, static_in, static_out, pool)

    # (kept for callers that look at the full-batch buffers)
    @property
    def static_in(self):
        return self._graphs[max(self._graphs)][1]

    @property
    def static_out(self):
        return self._graphs[max(self._graphs)][2]

    def __call__(self, batch: torch.Tensor) -> torch.Tensor:
        if tuple(batch.shape[1:]) != self._shape or batch.dtype != self._dtype:
            with torch.no_grad():
                return self.net(batch)
        b = batch.shape[0]
        if b not in self._graphs:
            self._capture(batch.clone())
        graph, static_in, static_out, _ = self._graphs[b]
        static_in.copy_(batch)
        graph.replay()
        return static_out


# ---- geometry -----------------------------------------------------------------------------------------
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


_IMPORTANCE_CACHE: Dict[tuple, torch.Tensor] = {}


def importance_map(roi_size: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125,
                   device="cpu") -> Optional[torch.Tensor]:
    """Cached per (roi, mode, sigma, device): building the 128^3 Gaussian on the host costs ~50 ms."""
    if mode == "constant":
        return None
    key = (tuple(int(r) for r in roi_size), mode, float(sigma_scale), str(device))
    hit = _IMPORTANCE_CACHE.get(key)
    if hit is None:
        hit = _IMPORTANCE_CACHE[key] = _importance_map(roi_size, mode, sigma_scale, device)
    return hit


def _importance_map(roi_size: Sequence[int], mode: str = "constant", sigma_scale: float = 0.125,
                    device="cpu") -> Optional[torch.Tensor]:
    """MONAI ``compute_importance_map``: ``None`` for "constant" (a weight of 1), else the product of per-axis
    Gaussians centred at ``roi // 2`` with sigma = ``sigma_scale * roi``, normalised to a maximum of 1 and with
    its zeros lifted to the smallest positive value (fp32, shape = roi)."""
    if mode == "constant":
        return None
    if mode != "gaussian":
        raise ValueError(f"mode must be 'constant' or 'gaussian', got {mode!r}")
    imp = torch.ones((), dtype=torch.float64)
    for k, r in enumerate(roi_size):
        x = torch.arange(r, dtype=torch.float64) - (r // 2)
        g = torch.exp(-0.5 * (x / (sigma_scale * r)) ** 2)
        shape = [1] * len(roi_size)
        shape[k] = r
        imp = imp * g.reshape(shape)
    imp = (imp / imp.max()).float()
    imp = torch.clamp(imp, min=float(imp[imp > 0].min()))
    return imp.contiguous().to(device)


def slab_bounds(extent: int, world: int) -> List[int]:
    """Borders of the ``world`` output slabs along the first spatial axis: slab r = [b[r], b[r+1])."""
    return [(extent * r) // world for r in range(world + 1)]


def window_runs(n_windows: int, world: int) -> List[int]:
    """Borders of the contiguous runs of the window list: rank r computes windows [b[r], b[r+1])."""
    return [(n_windows * r) // world for r in range(world + 1)]


@dataclass
class Piece:
    """Rows [s, e) (absolute first-axis coordinates) of window ``win``, computed by rank ``src``, accumulated by the
    owner ``dst`` of the slab they fall in; ``off`` = voxel offset inside the (src -> dst) transfer buffer."""
    win: int
    src: int
    dst: int
    s: int
    e: int
    off: int


@dataclass
class ShardPlan:
    dims: Tuple[int, int, int]
    roi: Tuple[int, int, int]
    wins: List[Tuple[int, int, int]]
    world: int
    bounds: List[int]
    runs: List[int]
    pieces: List[Piece] = field(default_factory=list)                 # global window order
    pair_voxels: Dict[Tuple[int, int], int] = field(default_factory=dict)  # (src, dst) -> voxels transferred

    def computed_by(self, win: int) -> int:
        for r in range(self.world):
            if self.runs[r] <= win < self.runs[r + 1]:
                return r
        raise IndexError(win)


def make_plan(dims: Sequence[int], roi: Sequence[int], overlap: float, world: int) -> ShardPlan:
    """Everything every rank needs to know about every transfer, from geometry alone (no negotiation)."""
    dims, roi = tuple(int(v) for v in dims), tuple(int(v) for v in roi)
    wins = window_list(dims, roi, overlap)
    plan = ShardPlan(dims, roi, wins, world, slab_bounds(dims[0], world), window_runs(len(wins), world))
    per_row = roi[1] * roi[2]
    for i, (a, _, _) in enumerate(wins):
        src = plan.computed_by(i)
        lo, hi = a, min(a + roi[0], dims[0])
        for dst in range(world):
            s, e = max(lo, plan.bounds[dst]), min(hi, plan.bounds[dst + 1])
            if s >= e:
                continue
            off = plan.pair_voxels.get((src, dst), 0)
            plan.pieces.append(Piece(i, src, dst, s, e, off))
            plan.pair_voxels[(src, dst)] = off + (e - s) * per_row
    return plan


# ---- device operations (the b200seg kernels; tests/CPU host-logic tests substitute a torch emulation) ---------
class _CudaOps:
    def check(self, x: torch.Tensor) -> None:
        if not x.is_cuda:
            raise RuntimeError("b200seg inference runs on CUDA tensors only")

    def accumulate(self, piece: torch.Tensor, imp: Optional[torch.Tensor], acc: torch.Tensor, cnt: torch.Tensor,
                   d0: int, h0: int, w0: int) -> None:
        """acc[d0:, h0:, w0:] += imp * piece; cnt += imp.  ``piece`` (rows, h, w, C) channels-last (ld >= C)."""
        lib = _lib.load()
        _, rows, wh, ww, c, ld = ops.cl_info(piece.unsqueeze(0))
        D, H, W, _ = acc.shape
        _lib.check(lib.b200seg_window_accumulate_weighted(
            ops.dtype_code(piece.dtype), piece.data_ptr(), ld, None if imp is None else imp.data_ptr(),
            acc.data_ptr(), cnt.data_ptr(), c, rows, wh, ww, D, H, W, d0, h0, w0,
            torch.cuda.current_stream().cuda_stream), "b200seg_window_accumulate_weighted")

    def argmax(self, acc: torch.Tensor, cnt: torch.Tensor, want_mean: bool):
        lib = _lib.load()
        D, H, W, c = acc.shape
        labels = torch.empty(D, H, W, dtype=torch.uint8, device=acc.device)
        mean = torch.empty_like(acc) if want_mean else None
        if D * H * W:
            _lib.check(lib.b200seg_accum_argmax(acc.data_ptr(), cnt.data_ptr(), labels.data_ptr(),
                                                None if mean is None else mean.data_ptr(), D * H * W, c,
                                                torch.cuda.current_stream().cuda_stream), "b200seg_accum_argmax")
        return labels, mean


# ---- the two halves of a rank's work (exposed so that tests can drive several emulated ranks in one process) ----
def _window_rows(cl: torch.Tensor, j: int, a: int, s: int, e: int) -> torch.Tensor:
    return cl[j, s - a:e - a]


class _RankState:
    """Rank ``rank``'s side of one sharded inference: ``compute()`` runs the forward passes of its windows and
    stages the row runs per destination; ``finish(recv)`` accumulates its slab in global window order and
    arg-maxes it."""

    def __init__(self, plan: ShardPlan, rank: int, x: torch.Tensor, predictor, sw_batch_size: int,
                 imp: Optional[torch.Tensor], dev_ops, n_classes: Optional[int], out_dtype: Optional[torch.dtype]):
        self.plan, self.rank, self.x, self.predictor, self.swb = plan, rank, x, predictor, sw_batch_size
        self.imp, self.ops = imp, dev_ops
        self.c, self.dtype = n_classes, out_dtype
        self.send: Dict[int, torch.Tensor] = {}
        b = plan.bounds
        self.slab = (b[rank], b[rank + 1])
        self.acc = self.cnt = None

    def _probe(self):
        """Class count / output dtype when the predictor does not state them: one ROI forward, no autograd."""
        r = self.plan.roi
        with torch.no_grad():
            y = self.predictor(self.x[:, :, :r[0], :r[1], :r[2]].contiguous())
        self.c, self.dtype = int(y.shape[1]), y.dtype

    def _alloc(self):
        if self.c is None or self.dtype is None:
            self._probe()
        dev = self.x.device
        _, H, W = self.plan.dims
        self.acc = torch.zeros(self.slab[1] - self.slab[0], H, W, self.c, dtype=torch.float32, device=dev)
        self.cnt = torch.zeros(self.slab[1] - self.slab[0], H, W, dtype=torch.float32, device=dev)
        if self.plan.world > 1:
            for (src, dst), nvox in self.plan.pair_voxels.items():
                if src == self.rank:
                    self.send[dst] = torch.empty(nvox * self.c, dtype=self.dtype, device=dev)

    def compute(self) -> Dict[int, torch.Tensor]:
        plan, r = self.plan, self.rank
        roi = plan.roi
        mine = list(range(plan.runs[r], plan.runs[r + 1]))
        by_win: Dict[int, List[Piece]] = {}
        for p in plan.pieces:
            if p.src == r:
                by_win.setdefault(p.win, []).append(p)
        for i in range(0, len(mine), self.swb):
            chunk = mine[i:i + self.swb]
            # a ragged last batch is filled up with repeats of its last window (their predictions are dropped): every
            # window is then computed at the SAME batch size whatever the sharding -- kernel grids and the grouping
            # of the InstanceNorm partial sums depend on the batch size, and bit-identical label maps across world
            # sizes need bit-identical window predictions
            padded = chunk + [chunk[-1]] * (self.swb - len(chunk))
            batch = torch.cat([self.x[:, :, a:a + roi[0], b:b + roi[1], c:c + roi[2]]
                               for a, b, c in (plan.wins[k] for k in padded)], 0)
            with torch.no_grad():
                pred = self.predictor(batch)
            cl = ops.to_channels_last(pred) if pred.is_cuda else pred.permute(0, 2, 3, 4, 1)
            if self.acc is None:
                if self.c is None:
                    self.c, self.dtype = int(cl.shape[-1]), cl.dtype
                self._alloc()
            for j, k in enumerate(chunk):
                a, b, c0 = plan.wins[k]
                for p in by_win[k]:
                    rows = _window_rows(cl, j, a, p.s, p.e)
                    if plan.world == 1:  # single rank: window order is the issue order, no staging copy
                        self.ops.accumulate(rows, None if self.imp is None else self.imp[p.s - a:p.e - a],
                                            self.acc, self.cnt, p.s - self.slab[0], b, c0)
                    else:
                        n = (p.e - p.s) * roi[1] * roi[2] * self.c
                        self.send[p.dst][p.off * self.c:p.off * self.c + n].view(p.e - p.s, roi[1], roi[2],
                                                                                   self.c).copy_(rows)
        if self.acc is None:  # no window for this rank (more ranks than windows): it still owns a slab
            self._alloc()
        return self.send

    def recv_buffers(self) -> Dict[int, torch.Tensor]:
        return {src: torch.empty(nvox * self.c, dtype=self.dtype, device=self.x.device)
                for (src, dst), nvox in self.plan.pair_voxels.items() if dst == self.rank and src != self.rank}

    def finish(self, recv: Dict[int, torch.Tensor], want_mean: bool):
        plan, r = self.plan, self.rank
        roi = plan.roi
        if plan.world > 1:
            for p in plan.pieces:  # global window order: the additions of the single-rank run, in its order
                if p.dst != r:
                    continue
                buf = self.send[r] if p.src == r else recv[p.src]
                n = (p.e - p.s) * roi[1] * roi[2] * self.c
                rows = buf[p.off * self.c:p.off * self.c + n].view(p.e - p.s, roi[1], roi[2], self.c)
                a, b, c0 = plan.wins[p.win]
                self.ops.accumulate(rows, None if self.imp is None else self.imp[p.s - a:p.e - a],
                                    self.acc, self.cnt, p.s - self.slab[0], b, c0)
        return self.ops.argmax(self.acc, self.cnt, want_mean)


def _exchange(state: _RankState, group=None) -> Dict[int, torch.Tensor]:
    """Point-to-point transfer of the staged row runs to the slab owners (sizes are known from the plan)."""
    recv = state.recv_buffers()
    reqs = []
    for src, buf in sorted(recv.items()):
        reqs.append(dist.P2POp(dist.irecv, buf, src, group))
    for dst, buf in sorted(state.send.items()):
        if dst != state.rank:
            reqs.append(dist.P2POp(dist.isend, buf, dst, group))
    if reqs:
        for w in dist.batch_isend_irecv(reqs):
            w.wait()
    return recv


def _gather_slabs(local: torch.Tensor, plan: ShardPlan, group=None) -> torch.Tensor:
    """All-gather of the per-rank slabs (first axis) into the whole volume; slabs are padded to the deepest one so
    that one equal-size collective does it."""
    depth = max(plan.bounds[r + 1] - plan.bounds[r] for r in range(plan.world))
    tail = tuple(local.shape[1:])
    if local.shape[0] < depth:
        local = torch.cat([local, local.new_zeros((depth - local.shape[0],) + tail)], 0)
    out = local.new_empty((plan.world * depth,) + tail)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    if plan.dims[0] == plan.world * depth:
        return out
    return torch.cat([out[r * depth:r * depth + plan.bounds[r + 1] - plan.bounds[r]] for r in range(plan.world)], 0)


def sliding_window_inference(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int,
                             predictor: Callable[[torch.Tensor], torch.Tensor], overlap: float = 0.25,
                             mode: str = "constant", sigma_scale: float = 0.125, return_logits: bool = False,
                             rank: Optional[int] = None, world: Optional[int] = None, group=None,
                             n_classes: Optional[int] = None, _dev_ops=None):
    """``inputs`` (1, Cin, D, H, W) -> uint8 label map (1, D, H, W) (and, with ``return_logits``, the averaged fp32
    logits (1, C, D, H, W)).  Under ``torch.distributed`` every rank passes the same volume and gets the whole
    label map; the windows and the output slabs are sharded as the module docstring describes."""
    if inputs.dim() != 5 or inputs.shape[0] != 1:
        raise ValueError("sliding_window_inference takes one 3-D volume: (1, C, D, H, W)")
    dev_ops = _dev_ops or _CudaOps()
    dev_ops.check(inputs)
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    dims = tuple(int(v) for v in inputs.shape[2:])
    roi = tuple(int(v) for v in roi_size)
    pad = []
    for sz, r in zip(reversed(dims), reversed(roi)):
        diff = max(r - sz, 0)
        pad += [diff // 2, diff - diff // 2]
    x = F.pad(inputs, pad) if any(pad) else inputs
    plan = make_plan(x.shape[2:], roi, overlap, world)
    imp = importance_map(roi, mode, sigma_scale, x.device)
    if n_classes is None:
        n_classes = getattr(predictor, "out_channels", None)
    out_dtype = getattr(predictor, "compute_dtype", None)
    state = _RankState(plan, rank, x, predictor, sw_batch_size, imp, dev_ops, n_classes, out_dtype)
    state.compute()
    recv = _exchange(state, group) if world > 1 else {}
    labels, mean = state.finish(recv, return_logits)
    if world > 1:
        labels = _gather_slabs(labels, plan, group)
        if return_logits:
            mean = _gather_slabs(mean, plan, group)
    sl = tuple(slice(max(r - s, 0) // 2, max(r - s, 0) // 2 + s) for s, r in zip(dims, roi))
    labels = labels[sl].unsqueeze(0)
    if return_logits:
        return labels, mean[sl].permute(3, 0, 1, 2).unsqueeze(0)
    return labels


def emulate_ranks(inputs: torch.Tensor, roi_size: Sequence[int], sw_batch_size: int, predictor, world: int,
                  overlap: float = 0.25, mode: str = "constant", sigma_scale: float = 0.125, _dev_ops=None):
    """Testing aid: the ``world``-rank algorithm executed rank after rank in ONE process on one device (the exchange
    becomes a hand-over of the staging buffers).  ``inputs`` must not need ROI padding.  Returns (labels (1, D, H, W)
    uint8, mean logits (1, C, D, H, W) fp32, plan)."""
    dev_ops = _dev_ops or _CudaOps()
    dev_ops.check(inputs)
    plan = make_plan(inputs.shape[2:], roi_size, overlap, world)
    imp = importance_map(plan.roi, mode, sigma_scale, inputs.device)
    states = [_RankState(plan, r, inputs, predictor, sw_batch_size, imp, dev_ops,
                         getattr(predictor, "out_channels", None), getattr(predictor, "compute_dtype", None))
              for r in range(world)]
    for st in states:
        st.compute()
    labs, means = [], []
    for r, st in enumerate(states):
        recv = {src: states[src].send[r] for src in range(world) if src != r and r in states[src].send}
        lab, mean = st.finish(recv, True)
        labs.append(lab)
        means.append(mean)
    return torch.cat(labs, 0).unsqueeze(0), torch.cat(means, 0).permute(3, 0, 1, 2).unsqueeze(0), plan
