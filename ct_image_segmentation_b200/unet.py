"""B200-native drop-in for ``monai.networks.nets.UNet`` as the reference uses it.

Reference seam: ``capstone/models/__init__.py:3`` re-exports ``monai.networks.nets.UNet``; it
is built at ``capstone/volumetric/base_trainer.py:65-72`` / ``capstone/training/base_trainer.py:72-79``
and called at ``capstone/volumetric/base_trainer.py:74-78``.

The module tree (names, parameter shapes, PyTorch weight layouts) is the MONAI-0.3 tree of
SURVEY.md Appendix A, so ``state_dict()`` / ``load_state_dict()`` / attribute paths such as
``unet.model[2][1].conv.unit0.conv`` (``capstone/interpretability.py:88``) keep working.  The leaf
``torch.nn`` modules only *hold* the parameters: the arithmetic never goes through them.
``forward`` runs an explicit forward plan on hand-written sm_100a kernels through the C ABI
(``include/b200seg.h``) and registers ONE autograd node whose backward runs the explicit reverse
plan (dgrad / wgrad / InstanceNorm+PReLU backward, in-place gradient fan-in, zero-copy skip
concatenation in both directions).
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import _lib, ops
from .ops import ConvGeom

# instances (voxels per sample) up to which the InstanceNorm+PReLU kernel reduces the convolution's partial
# statistics itself; above it the separate finalisation launch is cheaper than every block re-reading them
# (measured on cfg3, ms/step: never 2.291, <= 16^3 2.301, <= 32^3 2.312, <= 64^3 2.288, always 2.330)
_DEFER_STATS_MAX_VOX = int(os.environ.get("B200SEG_DEFER_STATS_MAX_VOX", str(64 ** 3)))

# dgrad fused with the reduction pass of the InstanceNorm + PReLU backward it feeds (ops.conv_dgrad_instnorm_partials)
_FUSE_IN_BWD = os.environ.get("B200SEG_FUSE_IN_BWD", "1") == "1"

_CONV = {2: nn.Conv2d, 3: nn.Conv3d}
_CONVT = {2: nn.ConvTranspose2d, 3: nn.ConvTranspose3d}
_INORM = {2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}


class Convolution(nn.Sequential):
    """Parameter holder for MONAI ``Convolution``: ``conv`` [-> ``norm`` -> ``act``] (A.2)."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3,
                 conv_only=False, is_transposed=False):
        super().__init__()
        pad = (kernel_size - 1) // 2
        if is_transposed:
            conv = _CONVT[dimensions](in_channels, out_channels, kernel_size, stride=strides,
                                      padding=pad, output_padding=strides - 1, bias=True)
        else:
            conv = _CONV[dimensions](in_channels, out_channels, kernel_size, stride=strides,
                                     padding=pad, bias=True)
        self.add_module("conv", conv)
        self.conv_only = conv_only
        if not conv_only:
            self.add_module("norm", _INORM[dimensions](out_channels))
            self.add_module("act", nn.PReLU())
        self.geom = ConvGeom(dimensions, in_channels, out_channels, kernel_size, strides, is_transposed)

    def forward(self, x):  # pragma: no cover - the engine never calls leaf modules
        raise RuntimeError("b200seg blocks are executed by UNet.forward, not called directly")


class ResidualUnit(nn.Module):
    """Parameter holder for MONAI ``ResidualUnit``: ``conv(x) + residual(x)`` (A.3)."""

    def __init__(self, dimensions, in_channels, out_channels, strides=1, kernel_size=3, subunits=2,
                 last_conv_only=False):
        super().__init__()
        self.conv = nn.Sequential()
        self.residual = nn.Identity()
        self.res_geom: Optional[ConvGeom] = None
        subunits = max(1, subunits)
        sc, ss = in_channels, strides
        for su in range(subunits):
            self.conv.add_module(
                f"unit{su:d}",
                Convolution(dimensions, sc, out_channels, strides=ss, kernel_size=kernel_size,
                            conv_only=last_conv_only and su == subunits - 1))
            sc, ss = out_channels, 1
        if strides != 1 or in_channels != out_channels:
            rk = kernel_size if strides != 1 else 1
            self.residual = _CONV[dimensions](in_channels, out_channels, rk, strides, (rk - 1) // 2,
                                              bias=True)
            self.res_geom = ConvGeom(dimensions, in_channels, out_channels, rk, strides, False)
        self.out_channels = out_channels

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("b200seg blocks are executed by UNet.forward, not called directly")


class SkipConnection(nn.Module):
    """``cat([x, submodule(x)], 1)`` -- realised zero-copy by the engine (A.3)."""

    def __init__(self, submodule):
        super().__init__()
        self.submodule = submodule

    def forward(self, x):  # pragma: no cover
        raise RuntimeError("b200seg blocks are executed by UNet.forward, not called directly")


class _Level(nn.Sequential):
    """(down, SkipConnection(sub), up) triple of one U-Net level."""


def _out_channels(layer) -> int:
    if isinstance(layer, ResidualUnit):
        return layer.out_channels
    if isinstance(layer, Convolution):
        return layer.geom.cout
    if isinstance(layer, _Level):
        return _out_channels(layer[2])
    return _out_channels(layer[-1])  # Sequential(conv, residual unit)


class UNet(nn.Module):
    """``UNet(dimensions, in_channels, out_channels, channels, strides, kernel_size=3,
    up_kernel_size=3, num_res_units=0, act="PRELU", norm="INSTANCE", dropout=0)``.

    Extra keyword (build-specific): ``dtype`` -- ``torch.bfloat16`` (production: bf16 storage,
    fp32 accumulation) or ``torch.float32`` (check mode).
    """

    def __init__(self, dimensions: int, in_channels: int, out_channels: int, channels: Sequence[int],
                 strides: Sequence[int], kernel_size=3, up_kernel_size=3, num_res_units: int = 0,
                 act="PRELU", norm="INSTANCE", dropout=0, dtype: torch.dtype = torch.bfloat16):
        super().__init__()
        if dimensions not in (2, 3):
            raise NotImplementedError("b200seg UNet supports dimensions 2 and 3")
        if kernel_size != 3 or up_kernel_size != 3:
            raise NotImplementedError("b200seg UNet supports kernel_size = up_kernel_size = 3")
        if str(act).upper() not in ("PRELU", "ACT.PRELU") or str(norm).upper() not in ("INSTANCE", "NORM.INSTANCE"):
            raise NotImplementedError("b200seg UNet supports act=PRELU, norm=INSTANCE")
        if dropout:
            raise NotImplementedError("b200seg UNet supports dropout=0")
        if len(channels) < 2 or len(strides) != len(channels) - 1:
            raise ValueError("need len(channels) >= 2 and len(strides) == len(channels) - 1")
        if any(s not in (1, 2) for s in strides):
            raise NotImplementedError("b200seg UNet supports strides 1 and 2")
        if dtype not in (torch.float32, torch.bfloat16):
            raise TypeError("dtype must be torch.float32 or torch.bfloat16")
        self.dimensions = dimensions
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.channels = list(channels)
        self.strides = list(strides)
        self.kernel_size = kernel_size
        self.up_kernel_size = up_kernel_size
        self.num_res_units = num_res_units
        self.compute_dtype = dtype
        self._packed: Dict = {}

        def level(inc, outc, chans, strds, is_top):
            c, s = chans[0], strds[0]
            if len(chans) > 2:
                sub, upc = level(c, c, chans[1:], strds[1:], False), 2 * c
            else:
                sub, upc = self._down(c, chans[1], 1), c + chans[1]
            return _Level(self._down(inc, c, s), SkipConnection(sub), self._up(upc, outc, s, is_top))

        self.model = level(in_channels, out_channels, self.channels, self.strides, True)

    # ---- tree construction ---------------------------------------------------------------
    def _down(self, inc, outc, s):
        if self.num_res_units > 0:
            return ResidualUnit(self.dimensions, inc, outc, strides=s, kernel_size=self.kernel_size,
                                subunits=self.num_res_units)
        return Convolution(self.dimensions, inc, outc, strides=s, kernel_size=self.kernel_size)

    def _up(self, inc, outc, s, is_top):
        conv = Convolution(self.dimensions, inc, outc, strides=s, kernel_size=self.up_kernel_size,
                           conv_only=is_top and self.num_res_units == 0, is_transposed=True)
        if self.num_res_units > 0:
            return nn.Sequential(conv, ResidualUnit(self.dimensions, outc, outc, strides=1,
                                                    kernel_size=self.kernel_size, subunits=1,
                                                    last_conv_only=is_top))
        return conv

    # ---- public forward ------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != self.dimensions + 2 or x.shape[1] != self.in_channels:
            raise ValueError(f"expected input (B, {self.in_channels}, *spatial[{self.dimensions}]), "
                             f"got {tuple(x.shape)}")
        if not x.is_cuda:
            raise RuntimeError("b200seg UNet runs on a CUDA (sm_100a) device only; there is no CPU path")
        div = 1
        for s in self.strides:
            div *= s
        if any(int(n) % div for n in x.shape[2:]):
            raise ValueError(f"spatial extents {tuple(x.shape[2:])} must be divisible by {div}")
        params = list(self.parameters())
        return _UNetFunction.apply(self, x, *params)

    def forward_debug(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        """Per-layer outputs as fp32 NC[D]HW tensors keyed by module path (parity tests)."""
        with torch.no_grad():
            saved: Dict = {}
            out = self._run_forward(ops.to_channels_last(x, self.compute_dtype), saved, keep_all=True)
        names = {m: n for n, m in self.named_modules()}
        taps = {"": ops.from_channels_last(out, self.dimensions).float()}
        for m, s in saved.items():
            if isinstance(m, Convolution):
                if s.get("c") is not None:
                    taps[names[m] + ".conv"] = ops.from_channels_last(s["c"], self.dimensions).float()
                if s.get("out") is not None and s.get("res") is None:  # (fused residual: not the module's own output)
                    taps[names[m]] = ops.from_channels_last(s["out"], self.dimensions).float()
        return taps

    # ---- weights ---------------------------------------------------------------------------
    def reset_packed_cache(self) -> None:
        """Forget the packed weights so that the next forward re-packs every layer (needed right
        before CUDA-graph capture: the pack kernels must be recorded into the graph)."""
        self._packed.clear()
        if getattr(self, "_batched", None) is not None:
            self._batched["state"] = None

    def _pack(self, param: torch.Tensor, geom: ConvGeom, kind: int, col: bool = False) -> torch.Tensor:
        if self.compute_dtype == torch.bfloat16:
            return self._batched["bufs"][(id(param), kind, col)]  # refreshed by _repack_all()
        key = (id(param), kind, self.compute_dtype)
        hit = self._packed.get(key)
        ver = (param._version, param.data_ptr())
        if hit is None or hit[0] != ver:
            hit = (ver, ops.pack_weight(geom, kind, param, self.compute_dtype))
            self._packed[key] = hit
        return hit[1]

    def _conv_params(self):
        """(conv module, geometry) of every convolution in the tree."""
        out = []
        for m in self.modules():
            if isinstance(m, Convolution):
                out.append((m.conv, m.geom))
            elif isinstance(m, ResidualUnit) and m.res_geom is not None:
                out.append((m.residual, m.res_geom))
        return out

    def _repack_all(self) -> None:
        """bf16 mode: repack EVERY weight (fprop + dgrad layouts) with one kernel launch whenever a
        parameter changed (optimiser step) -- the launch is recorded into the CUDA graph."""
        convs = self._conv_params()
        state = tuple((c.weight._version, c.weight.data_ptr()) for c, _ in convs)
        b = getattr(self, "_batched", None)
        if b is not None and b["ptrs"] == tuple(s[1] for s in state):
            if b["state"] == state:
                return
        else:  # (re)build the persistent buffers and the device tables
            # two tables: the first down layer's weights (needed at once, a few KB) and everything else
            # (95 % of the bytes; with the second stream enabled it is packed beside the first layers)
            early_ids = {id(p) for p in self.model[0].parameters()}
            bufs, entries, late = {}, [], []
            for conv, g in convs:
                entries_all = entries
                entries = entries_all if id(conv.weight) in early_ids else late
                for kind in ((_lib.W_CONVTR_FPROP, _lib.W_CONVTR_DGRAD) if g.transposed
                             else (_lib.W_CONV_FPROP, _lib.W_CONV_DGRAD)):
                    nbytes, tc_off = ops.packed_weight_layout(g, kind, self.compute_dtype)
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=conv.weight.device)
                    bufs[(id(conv.weight), kind, False)] = buf
                    k = g.kernel
                    # stride-1 layers with 16-aligned channel counts never leave the tcgen05 kernels in
                    # bf16 mode: their generic (CUDA-core) layout is not packed
                    tc_only = (g.stride == 1 and g.cin % 16 == 0 and g.cout % 16 == 0)
                    entries.append((conv.weight.data_ptr(), buf.data_ptr(), tc_off,
                                    k ** self.dimensions, g.cin, g.cout,
                                    kind | (_lib.PACK_TC_ONLY if tc_only else 0)))
                g1 = ops.col_geom(g)
                if g1 is not None:
                    # small-Cin layer run as a 1x1x1 conv on its im2col buffer: the PyTorch weight
                    # (cout, cin, taps) is the (cout, cin*taps) matrix of that conv as it stands
                    nbytes, tc_off = ops.packed_weight_layout(g1, _lib.W_CONV_FPROP, self.compute_dtype)
                    buf = torch.empty(nbytes, dtype=torch.uint8, device=conv.weight.device)
                    bufs[(id(conv.weight), _lib.W_CONV_FPROP, True)] = buf
                    entries.append((conv.weight.data_ptr(), buf.data_ptr(), tc_off, 1, g1.cin, g1.cout,
                                    _lib.W_CONV_FPROP))
                entries = entries_all
            dev = convs[0][0].weight.device
            b = {"bufs": bufs, "table": ops.make_pack_table(entries, dev), "n": len(entries),
                 "table_late": ops.make_pack_table(late, dev) if late else None, "n_late": len(late),
                 "ptrs": tuple(s[1] for s in state)}
            self._batched = b
        ops.pack_weights_batched(b["table"], b["n"])
        if b["n_late"]:
            side = self._wgrad_side if self.wgrad_stream else None
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    ops.pack_weights_batched(b["table_late"], b["n_late"])
                self._pack_join = True  # joined after the first down layer (_fwd_level)
            else:
                ops.pack_weights_batched(b["table_late"], b["n_late"])
        b["state"] = state

    def _w_fprop(self, conv: nn.Module, geom: ConvGeom):
        kind = _lib.W_CONVTR_FPROP if geom.transposed else _lib.W_CONV_FPROP
        return self._pack(conv.weight, geom, kind)

    def _w_dgrad(self, conv: nn.Module, geom: ConvGeom):
        kind = _lib.W_CONVTR_DGRAD if geom.transposed else _lib.W_CONV_DGRAD
        return self._pack(conv.weight, geom, kind)

    # ---- forward plan ----------------------------------------------------------------------
    def _new(self, like: torch.Tensor, spatial, c: int) -> torch.Tensor:
        return ops.alloc_activation(like.shape[0], tuple(spatial), c, self.compute_dtype, like.device)

    def _run_forward(self, x_cl: torch.Tensor, saved: Dict, keep_all: bool = False,
                     use_cols: bool = True) -> torch.Tensor:
        if self.compute_dtype == torch.bfloat16 and x_cl.is_cuda:
            self._repack_all()
        # small-Cin first layers as im2col + 1x1x1 tcgen05 convs (bf16; no input gradient possible there)
        self._cols = {} if (use_cols and self.compute_dtype == torch.bfloat16) else None
        try:
            return self._fwd_level(self.model, x_cl, saved, None, keep_all)
        finally:
            self._cols = None

    def _col_input(self, g: ConvGeom, x: torch.Tensor):
        """(1x1x1 geometry, im2col buffer) for a small-Cin layer in col mode, else (None, None).  Layers
        with the same input and geometry (unit0 conv and residual conv of a ResidualUnit) share it."""
        if getattr(self, "_cols", None) is None:
            return None, None
        g1 = ops.col_geom(g)
        if g1 is None:
            return None, None
        key = (x.data_ptr(), tuple(x.shape), g.kernel, g.stride)
        col = self._cols.get(key)
        if col is None:
            col = ops.im2col(g, x)
            self._cols[key] = col
        return g1, col

    def _fwd_level(self, lvl: _Level, x, saved, dst, keep):
        down, skip, up = lvl[0], lvl[1], lvl[2]
        sub = skip.submodule
        c_x, c_sub = _out_channels(down), _out_channels(sub)
        sp = self._down_geom(down).out_spatial(*x.shape[1:4])
        cat = self._new(x, sp, c_x + c_sub)  # skip concatenation buffer: [x | sub(x)]
        xd = self._fwd_layer(down, x, saved, cat[..., :c_x], keep)
        if getattr(self, "_pack_join", False):
            torch.cuda.current_stream().wait_stream(self._wgrad_side)  # the other layers' weights are packed
            self._pack_join = False
        if isinstance(sub, _Level):
            self._fwd_level(sub, xd, saved, cat[..., c_x:], keep)
        else:
            self._fwd_layer(sub, xd, saved, cat[..., c_x:], keep)
        saved[lvl] = {"c_x": c_x}
        return self._fwd_layer(up, cat, saved, dst, keep)

    @staticmethod
    def _down_geom(down) -> ConvGeom:
        return down.conv.unit0.geom if isinstance(down, ResidualUnit) else down.geom

    def _fwd_layer(self, layer, x, saved, dst, keep):
        if isinstance(layer, ResidualUnit):
            return self._fwd_resunit(layer, x, saved, dst, keep)
        if isinstance(layer, Convolution):
            return self._fwd_convolution(layer, x, saved, dst, None, keep)
        h = self._fwd_convolution(layer[0], x, saved, None, None, keep)  # up: ConvTranspose block
        return self._fwd_resunit(layer[1], h, saved, dst, keep)

    def _fwd_convolution(self, m: Convolution, x, saved, dst, residual, keep):
        g = m.geom
        sp = g.out_spatial(*x.shape[1:4])
        g1, col = self._col_input(g, x)
        if g1 is not None:
            g, x = g1, col
            wp = self._pack(m.conv.weight, g1, _lib.W_CONV_FPROP, col=True)
        else:
            wp = self._w_fprop(m.conv, g)
        bias = m.conv.bias.detach()
        if m.conv_only:
            y = dst if dst is not None else self._new(x, sp, g.cout)
            ops.conv_fprop(g, x, wp, bias, y, residual)
            # with a fused residual the module's own output is never materialised: no tap
            saved[m] = {"x": x, "c": None, "out": y if keep else None, "res": residual if keep else None,
                        "col_geom": g1}
            return y
        c = self._new(x, sp, g.cout)
        a = dst if dst is not None else self._new(x, sp, g.cout)
        # InstanceNorm statistics come out of the convolution's epilogue as per-CTA partials where the
        # kernel supports it; for instances of up to _DEFER_STATS_MAX_VOX voxels the InstanceNorm+PReLU
        # kernel finalises them itself (one launch less on the serial chain)
        alpha = m.act.weight.detach()
        defer = c.is_cuda and g.cout <= 256 and sp[0] * sp[1] * sp[2] <= _DEFER_STATS_MAX_VOX
        handle = ops.conv_fprop_partials(g, x, wp, bias, c) if defer else None
        if handle is not None:
            mean, rstd = ops.instnorm_prelu_fwd_partials(c, handle, alpha, a, residual, m.norm.eps)
        else:
            if defer:  # the layer ran on a kernel without the fusion: c is complete
                mean, rstd = ops.instnorm_stats(c, m.norm.eps)
            else:
                mean, rstd = ops.conv_fprop_stats(g, x, wp, bias, c, m.norm.eps)
            ops.instnorm_prelu_fwd(c, mean, rstd, alpha, a, residual, m.norm.eps)
        # tests (keep): "out" is what the kernel wrote -- with a fused residual that is module output + "res"
        saved[m] = {"x": x, "c": c, "mean": mean, "rstd": rstd, "out": a if keep else None,
                    "res": residual if keep else None, "col_geom": g1}
        return a

    def _fwd_resunit(self, ru: ResidualUnit, x, saved, dst, keep):
        units = list(ru.conv.children())
        if ru.res_geom is not None:
            rg = ru.res_geom
            r = self._new(x, rg.out_spatial(*x.shape[1:4]), rg.cout)
            rg1, rcol = self._col_input(rg, x)
            # the residual conv is only consumed by the unit's last InstanceNorm+PReLU: with the second
            # stream enabled it runs beside the unit0 -> unit1 chain
            side = self._wgrad_side if (self.wgrad_stream and x.is_cuda) else None
            if side is not None:
                side.wait_stream(torch.cuda.current_stream())
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                if rg1 is not None:
                    ops.conv_fprop(rg1, rcol, self._pack(ru.residual.weight, rg1, _lib.W_CONV_FPROP, col=True),
                                   ru.residual.bias.detach(), r)
                else:
                    ops.conv_fprop(rg, x, self._w_fprop(ru.residual, rg), ru.residual.bias.detach(), r)
        else:
            r = x
            rg1 = rcol = None
            side = None
        h = x
        for i, u in enumerate(units):
            last = i == len(units) - 1
            if last and side is not None:
                torch.cuda.current_stream().wait_stream(side)  # join before r is read
            h = self._fwd_convolution(u, h, saved, dst if last else None, r if last else None, keep)
        saved[ru] = {"x": x, "col_geom": rg1, "col": rcol, "r": r if (keep and ru.res_geom is not None) else None}
        return h

    # ---- backward plan -------------------------------------------------------------------------
    def _run_backward(self, saved: Dict, g_out: torch.Tensor, need_gx: bool, taps: Optional[Dict] = None):
        """Reverse plan.  ``taps`` (tests only) receives the local inputs/outputs of every
        Convolution's backward (g_out, g_c, x, c, mean, rstd)."""
        grads: Dict[torch.Tensor, torch.Tensor] = {}
        self._bwd_taps = taps
        self._wgrad_keep = []
        self._in_partials = {}  # Convolution -> (handle, g_out): InstanceNorm-backward sums left by the producing dgrad
        gx = self._bwd_level(self.model, g_out, saved, grads, need_gx, None, False, 0)
        side = self._wgrad_side if self.wgrad_stream else None
        if side is not None and g_out.is_cuda:
            torch.cuda.current_stream().wait_stream(side)  # join: weight gradients complete
        self._wgrad_keep = []
        return grads, gx

    # Weight gradients are leaves of the backward data flow (nothing in the step reads them before the
    # optimiser), while dgrad -> InstanceNorm backward -> dgrad is a serial chain whose deep layers fill a
    # fraction of the 148 SMs.  With ``wgrad_stream`` the wgrad launches go to a second stream (a parallel
    # branch of the captured graph) and overlap that chain.  Operands stay referenced until the join:
    # the caching allocator must not hand their memory to later main-stream kernels.
    wgrad_stream = False
    _wgrad_side = None

    def enable_wgrad_stream(self, on: bool = True) -> None:
        self.wgrad_stream = bool(on)
        if on and self._wgrad_side is None:
            self._wgrad_side = torch.cuda.Stream()

    # Gradient sink (GraphedTrainStep): {parameter: contiguous fp32 view of a flat gradient bucket}.  With a
    # sink the weight-gradient / bias / PReLU kernels write straight into the bucket, dead biases are left to
    # the bucket's zeros, and the autograd node reports no parameter gradients (nothing to gather afterwards).
    _grad_sink: Optional[Dict] = None
    # called as hook(depth) from the reverse plan when every launch of level `depth` and below has been
    # issued (main stream + weight-gradient stream): the point where their bucket range can be all-reduced
    _level_done_hook = None

    def bind_grad_sink(self, sink: Optional[Dict], level_done_hook=None) -> None:
        self._grad_sink = sink
        self._level_done_hook = level_done_hook

    def _wgrad(self, geom: ConvGeom, x, g, want_bias: bool = True, weight=None, bias=None):
        sink = self._grad_sink
        out_w = out_b = None
        if sink is not None:  # (a frozen parameter is not in the bucket: its gradient goes to a scratch tensor)
            out_w = sink.get(weight)
            out_b = sink.get(bias) if want_bias else None
        if not (self.wgrad_stream and x.is_cuda):
            return ops.conv_wgrad(geom, x, g, want_bias=want_bias, out_w=out_w, out_b=out_b)
        side = self._wgrad_side
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            gw, gb = ops.conv_wgrad(geom, x, g, want_bias=want_bias, out_w=out_w, out_b=out_b)
        self._wgrad_keep.append((x, g, gw, gb))
        return gw, gb

    def _bwd_level(self, lvl: _Level, g_out, saved, grads, need_gx, gx_dst, gx_accum, depth: int = 0):
        down, skip, up = lvl[0], lvl[1], lvl[2]
        sub = skip.submodule
        c_x = saved.pop(lvl)["c_x"]
        g_cat = self._bwd_layer(up, g_out, saved, grads, True, None, False)
        g_x, g_sub = g_cat[..., :c_x], g_cat[..., c_x:]
        # the sub-network's input gradient is accumulated in place into the skip half of g_cat
        if isinstance(sub, _Level):
            self._bwd_level(sub, g_sub, saved, grads, True, g_x, True, depth + 1)
        else:
            self._bwd_layer(sub, g_sub, saved, grads, True, g_x, True)
        if self._level_done_hook is not None:
            self._level_done_hook(depth + 1)
        return self._bwd_layer(down, g_x, saved, grads, need_gx, gx_dst, gx_accum)

    def _bwd_layer(self, layer, g_out, saved, grads, need_gx, gx_dst, gx_accum):
        if isinstance(layer, ResidualUnit):
            return self._bwd_resunit(layer, g_out, saved, grads, need_gx, gx_dst, gx_accum)
        if isinstance(layer, Convolution):
            return self._bwd_convolution(layer, g_out, saved, grads, need_gx, gx_dst, gx_accum, None)
        # up layer = (ConvTranspose block, residual unit): the unit's input gradient is the block's output gradient
        g_h = self._bwd_resunit(layer[1], g_out, saved, grads, True, None, False, next_m=layer[0])
        return self._bwd_convolution(layer[0], g_h, saved, grads, need_gx, gx_dst, gx_accum, None)

    def _bwd_convolution(self, m: Convolution, g_out, saved, grads, need_gx, gx_dst, gx_accum,
                         gx_residual, next_m: Optional[Convolution] = None):
        """``next_m``: the Convolution (conv -> InstanceNorm -> PReLU) whose OUTPUT gradient this layer's input
        gradient is, when nothing else is added to it afterwards: its InstanceNorm backward starts with sums over
        exactly the tensor the dgrad epilogue holds in registers, so the dgrad leaves them as per-CTA partials
        (``ops.conv_dgrad_instnorm_partials``) and ``next_m``'s backward skips its reduction pass."""
        s = saved.pop(m)
        g, x = m.geom, s["x"]
        if m.conv_only:
            g_c = g_out
        else:
            c = s["c"]
            g_c = ops.alloc_like(c)
            sink = self._grad_sink
            out_da = None if sink is None else sink.get(m.act.weight)
            fused = getattr(self, "_in_partials", {}).pop(m, None)
            if fused is not None and fused[1] is g_out:
                dalpha = ops.instnorm_prelu_bwd_from_partials(c, s["mean"], s["rstd"], m.act.weight.detach(), g_out, g_c,
                                                              fused[0], m.norm.eps, out_dalpha=out_da)
            else:
                dalpha = ops.instnorm_prelu_bwd(c, s["mean"], s["rstd"], m.act.weight.detach(), g_out, g_c, m.norm.eps,
                                                out_dalpha=out_da)
            if sink is None:
                grads[m.act.weight] = dalpha
        if getattr(self, "_bwd_taps", None) is not None:
            self._bwd_taps[m] = {"g_out": g_out, "g_c": g_c, "x": x, "c": s.get("c"),
                                 "mean": s.get("mean"), "rstd": s.get("rstd")}
        # A bias in front of an affine-less InstanceNorm is cancelled by the mean subtraction: its
        # gradient is identically zero (sum_v g_c = 0 analytically; the reference holds rounding noise
        # there).  Only live biases (conv-only head) get a column sum.
        if s.get("col_geom") is not None:  # x is the im2col buffer; gw comes out as the (cout, cin*taps) matrix
            if need_gx:
                raise RuntimeError("input gradient requested through an im2col first layer")
            gw, gb = self._wgrad(s["col_geom"], x, g_c, m.conv_only, m.conv.weight, m.conv.bias)
            gw = gw.view(m.conv.weight.shape)
        else:
            gw, gb = self._wgrad(g, x, g_c, m.conv_only, m.conv.weight, m.conv.bias)
        if self._grad_sink is None:
            grads[m.conv.weight] = gw
            grads[m.conv.bias] = gb if gb is not None else torch.zeros_like(m.conv.bias)
        if not need_gx:
            return None
        gx = gx_dst if gx_dst is not None else ops.alloc_like(x)
        handle = None
        if next_m is not None and _FUSE_IN_BWD and not next_m.conv_only and x.is_cuda:
            sn = saved.get(next_m)
            if sn is not None and sn.get("c") is not None and tuple(sn["c"].shape) == tuple(gx.shape):
                handle = ops.conv_dgrad_instnorm_partials(
                    g, g_c, self._w_dgrad(m.conv, g), gx, sn["c"], sn["mean"], sn["rstd"], next_m.act.weight.detach(),
                    residual=gx_residual, flags=_lib.CONV_ACCUMULATE if gx_accum else 0)
        if handle is None:
            ops.conv_dgrad(g, g_c, self._w_dgrad(m.conv, g), gx, residual=gx_residual, accumulate=gx_accum)
        else:
            self._in_partials[next_m] = (handle, gx)
        return gx

    def _bwd_resunit(self, ru: ResidualUnit, g_out, saved, grads, need_gx, gx_dst, gx_accum,
                     next_m: Optional[Convolution] = None):
        units = list(ru.conv.children())
        sru = saved.pop(ru)
        x = sru["x"]
        g = g_out
        for i in range(len(units) - 1, 0, -1):
            # unit i's input gradient is unit i-1's output gradient, and nothing else is added to it
            g = self._bwd_convolution(units[i], g, saved, grads, True, None, False, None, next_m=units[i - 1])
        if ru.res_geom is None:
            # identity residual: d/dx = dgrad(unit0) + g_out, fused as the dgrad epilogue addend (the sum is final
            # there: a following block's InstanceNorm backward can take its sums from that epilogue)
            return self._bwd_convolution(units[0], g, saved, grads, need_gx, gx_dst, gx_accum, g_out,
                                         next_m=next_m if (gx_dst is None and not gx_accum) else None)
        gx = self._bwd_convolution(units[0], g, saved, grads, need_gx, gx_dst, gx_accum, None)
        rg = ru.res_geom
        if getattr(self, "_bwd_taps", None) is not None:
            self._bwd_taps[ru] = {"g_out": g_out, "x": x, "col": sru.get("col"), "col_geom": sru.get("col_geom"),
                                  "r": sru.get("r")}
        if sru.get("col_geom") is not None:
            gw, gb = self._wgrad(sru["col_geom"], sru["col"], g_out, True, ru.residual.weight, ru.residual.bias)
            gw = gw.view(ru.residual.weight.shape)
        else:
            gw, gb = self._wgrad(rg, x, g_out, True, ru.residual.weight, ru.residual.bias)
        if self._grad_sink is None:
            grads[ru.residual.weight], grads[ru.residual.bias] = gw, gb
        if need_gx:
            ops.conv_dgrad(rg, g_out, self._w_dgrad(ru.residual, rg), gx, accumulate=True)
        return gx


class _UNetFunction(torch.autograd.Function):
    """One autograd node for the whole network: explicit forward and reverse plans."""

    @staticmethod
    def forward(ctx, net: UNet, x: torch.Tensor, *params):
        x_cl = ops.to_channels_last(x.detach(), net.compute_dtype)
        saved: Dict = {}
        out = net._run_forward(x_cl, saved, use_cols=not ctx.needs_input_grad[1])
        need = any(ctx.needs_input_grad[1:])
        ctx.net = net
        ctx.saved = saved if need else None
        ctx.params = params
        ctx.x_dtype = x.dtype
        return ops.from_channels_last(out, net.dimensions)

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        net = ctx.net
        if ctx.saved is None:
            raise RuntimeError("UNet backward called but the forward did not record activations")
        g_cl = ops.to_channels_last(g, net.compute_dtype)
        saved, ctx.saved = ctx.saved, None
        grads, gx = net._run_backward(saved, g_cl, ctx.needs_input_grad[1])
        gx_out = None
        if gx is not None:
            gx_out = ops.from_channels_last(gx, net.dimensions).to(ctx.x_dtype)
        return (None, gx_out, *[grads.get(p) for p in ctx.params])
