"""Thin host-side wrappers: torch tensors in, C-ABI calls on the current CUDA stream.

Activations are *channels-last* tensors of shape ``(N, D, H, W, C)`` (2-D data: ``D == 1``)
whose last-dim stride is 1 and whose voxel stride ``ld >= C`` is uniform, i.e. either a
contiguous tensor or a channel slice ``buf[..., c0:c1]`` of one (zero-copy concatenation).
Nothing here falls back to PyTorch arithmetic: a missing library or a failed launch raises.
"""
from __future__ import annotations

import ctypes as C
import math
import weakref
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import BF16, F32, ConvDesc, DiceDesc, NormDesc

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"b200seg kernels take float32 or bfloat16 activations, got {dt}") from None


def cl_info(t: torch.Tensor) -> Tuple[int, int, int, int, int, int]:
    """(n, d, h, w, c, ld) of a channels-last activation; raises if the layout is not usable."""
    if t.dim() != 5:
        raise ValueError(f"expected (N, D, H, W, C), got shape {tuple(t.shape)}")
    n, d, h, w, c = t.shape
    sn, sd, sh, sw, sc = t.stride()
    ld = sw if w > 1 else (sh if h > 1 else (sd if d > 1 else (sn if n > 1 else c)))
    ok = (sc == 1 or c == 1) and ld >= c
    ok = ok and (w == 1 or sw == ld) and (h == 1 or sh == w * ld) and (d == 1 or sd == h * w * ld)
    ok = ok and (n == 1 or sn == d * h * w * ld)
    if not ok:
        raise ValueError(f"tensor is not channels-last with a uniform voxel stride: shape "
                         f"{tuple(t.shape)}, strides {t.stride()}")
    if not t.is_cuda:
        raise RuntimeError("b200seg ops need CUDA tensors (no CPU fallback)")
    return n, d, h, w, c, ld


def to_channels_last(x: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """(N, C, *S) logical tensor -> (N, D, H, W, C) channels-last (a view when already so)."""
    if x.dim() == 4:
        x = x.unsqueeze(2)
    y = x.permute(0, 2, 3, 4, 1)
    if dtype is not None and y.dtype != dtype:
        y = y.to(dtype)
    if y.is_contiguous():
        return y
    try:
        cl_info(y)  # uniform voxel stride (e.g. a zero-padded buffer): usable without a copy
        return y
    except (ValueError, RuntimeError):
        return y.contiguous()


def from_channels_last(y: torch.Tensor, dims: int) -> torch.Tensor:
    """(N, D, H, W, C) -> logical (N, C, D, H, W) / (N, C, H, W) view (no copy)."""
    out = y.permute(0, 4, 1, 2, 3)
    return out.squeeze(2) if dims == 2 else out


# ----------------------------------------------------------------------------------------------
# workspace (owned by the caller side of the ABI: a grow-only per-stream torch buffer)
# ----------------------------------------------------------------------------------------------
_workspaces = {}


def workspace(nbytes: int, device: torch.device) -> torch.Tensor:
    key = (device.index if device.index is not None else torch.cuda.current_device(), _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


# ----------------------------------------------------------------------------------------------
# convolution
# ----------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class ConvGeom:
    """One Conv / ConvTranspose layer of the U-Net (kernel 1 or 3, stride 1 or 2, same padding)."""
    dims: int          # 2 or 3
    cin: int
    cout: int
    kernel: int        # 1 or 3
    stride: int        # 1 or 2
    transposed: bool

    def out_spatial(self, d: int, h: int, w: int) -> Tuple[int, int, int]:
        k, s, p = self.kernel, self.stride, (self.kernel - 1) // 2
        if self.transposed:
            f = lambda n: (n - 1) * s - 2 * p + k + (s - 1)
        else:
            f = lambda n: (n + 2 * p - k) // s + 1
        return (d if self.dims == 2 else f(d)), f(h), f(w)

    def desc(self, n, in_sp, out_sp, x_ld, y_ld, r_ld, dtype, flags=0) -> ConvDesc:
        k, s, p = self.kernel, self.stride, (self.kernel - 1) // 2
        kd, sd, pd = (1, 1, 0) if self.dims == 2 else (k, s, p)
        return ConvDesc(n, self.cin, self.cout, *in_sp, *out_sp, kd, k, k, sd, s, s, pd, p, p,
                        x_ld, y_ld, r_ld, dtype_code(dtype), flags)


def pack_weight(geom: ConvGeom, kind: int, w: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """fp32 PyTorch-layout parameter -> kernel layout for one data-path use."""
    lib = _lib.load()
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    d = geom.desc(1, (1, 1, 1), (1, 1, 1), geom.cin, geom.cout, 0, dtype)
    nbytes = lib.b200seg_packed_weight_bytes(C.byref(d), kind)
    out = torch.empty(nbytes, dtype=torch.uint8, device=w.device)
    _lib.check(lib.b200seg_pack_weight(C.byref(d), kind, w.data_ptr(), out.data_ptr(), _stream()),
               "b200seg_pack_weight")
    return out


def packed_weight_layout(geom: ConvGeom, kind: int, dtype: torch.dtype):
    """(bytes of the packed buffer, byte offset of the tcgen05 layout inside it)."""
    lib = _lib.load()
    d = geom.desc(1, (1, 1, 1), (1, 1, 1), geom.cin, geom.cout, 0, dtype)
    return lib.b200seg_packed_weight_bytes(C.byref(d), kind), lib.b200seg_packed_weight_tc_offset(C.byref(d), kind)


def make_pack_table(entries, device) -> torch.Tensor:
    """Device-resident array of b200seg_pack_entry for pack_weights_batched."""
    arr = (_lib.PackEntry * len(entries))(*[_lib.PackEntry(*e) for e in entries])
    host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return host.to(device)


def pack_weights_batched(table: torch.Tensor, n: int) -> None:
    lib = _lib.load()
    _lib.check(lib.b200seg_pack_weights_batched(table.data_ptr(), n, _stream()), "b200seg_pack_weights_batched")


_PADDED = set()  # (storage pointer, C) of live zero-padded buffers handed out by alloc_activation


# Persistent pool of zero-padded buffers (CUDA-graph mode): every kernel keeps the channel padding
# zero, so a buffer zeroed ONCE can be handed out again every step instead of being memset
# (134 MB per 10-class tensor at 128^3 x 2).  Only valid when a step's activations are dead before
# the next step starts -- GraphedTrainStep switches it on; eager use keeps fresh torch.zeros.
_PAD_POOL = None   # None = off; else {"seq": int, "bufs": {(key, seq): tensor}}


class padded_buffer_pool:
    """Context manager: inside it, the k-th padded allocation reuses the k-th buffer of ``pool``
    (a dict owned by the caller, e.g. one per GraphedTrainStep)."""

    def __init__(self, pool: dict):
        self.pool = pool

    def __enter__(self):
        global _PAD_POOL
        self.prev = _PAD_POOL
        self.pool["seq"] = 0
        self.pool.setdefault("bufs", {})
        _PAD_POOL = self.pool
        return self

    def __exit__(self, *exc):
        global _PAD_POOL
        _PAD_POOL = self.prev
        return False


def alloc_activation(n: int, spatial, c: int, dtype: torch.dtype, device) -> torch.Tensor:
    """(N, *spatial, C) activation buffer.  For bf16 and C not a multiple of 16 (the 10-class
    layers) the buffer is zero-padded to the next multiple of 16 channels and marked, so the
    tcgen05 kernels can take it (B200SEG_CONV_PADDED_CHANNELS)."""
    if dtype == torch.bfloat16 and c % 16 != 0 and c >= 8:
        cp = (c + 15) // 16 * 16
        if _PAD_POOL is not None:
            key = ((n, *spatial, cp), str(device), _PAD_POOL["seq"])
            _PAD_POOL["seq"] += 1
            full = _PAD_POOL["bufs"].get(key)
            if full is None:
                full = torch.zeros((n, *spatial, cp), dtype=dtype, device=device)
                _PAD_POOL["bufs"][key] = full
                _PADDED.add((full.untyped_storage().data_ptr(), c))
            return full[..., :c]
        full = torch.zeros((n, *spatial, cp), dtype=dtype, device=device)
        key = (full.untyped_storage().data_ptr(), c)
        _PADDED.add(key)
        weakref.finalize(full, _PADDED.discard, key)  # views keep `full` alive through ._base
        return full[..., :c]
    return torch.empty((n, *spatial, c), dtype=dtype, device=device)


def alloc_like(t: torch.Tensor) -> torch.Tensor:
    return alloc_activation(t.shape[0], tuple(t.shape[1:4]), t.shape[4], t.dtype, t.device)


def _pad_safe(t: Optional[torch.Tensor]) -> bool:
    """True if the tcgen05 kernels may treat channels [C, round_up(C, 16)) of ``t`` as its own zero
    padding: C is a multiple of 16, or ``t`` views a live buffer that ``alloc_activation`` created
    zero-padded for exactly C channels (every kernel keeps that padding zero).  A channel slice of
    a concatenation buffer never qualifies."""
    if t is None or t.shape[-1] % 16 == 0:
        return True
    n, d, h, w, c = t.shape
    cp = (c + 15) // 16 * 16
    try:
        ld = cl_info(t)[5]
    except (ValueError, RuntimeError):
        return False
    return (ld == cp and t.storage_offset() == 0
            and (t.untyped_storage().data_ptr(), c) in _PADDED)


def _conv_call(geom: ConvGeom, src, wp, bias, residual, dst, data_grad: bool, flags: int):
    """Shared driver of the four data-path entry points.  ``src``/``dst`` are x/y for fprop and
    dy/dx for dgrad; the descriptor is always written in the layer's own terms."""
    lib = _lib.load()
    n, sd_, sh_, sw_, sc, s_ld = cl_info(src)
    n2, dd_, dh_, dw_, dc, d_ld = cl_info(dst)
    if n != n2 or src.dtype != dst.dtype:
        raise ValueError("conv: batch/dtype mismatch between source and destination")
    r_ld = 0
    if residual is not None:
        rn, rd, rh, rw, rc, r_ld = cl_info(residual)
        if (rn, rd, rh, rw, rc) != (n2, dd_, dh_, dw_, dc) or residual.dtype != dst.dtype:
            raise ValueError("conv: residual must match the destination")
    if not data_grad:
        in_sp, out_sp, x_ld, y_ld = (sd_, sh_, sw_), (dd_, dh_, dw_), s_ld, d_ld
        if (sc, dc) != (geom.cin, geom.cout):
            raise ValueError(f"conv fprop: channels {(sc, dc)} != {(geom.cin, geom.cout)}")
    else:
        in_sp, out_sp, x_ld, y_ld = (dd_, dh_, dw_), (sd_, sh_, sw_), d_ld, s_ld
        if (sc, dc) != (geom.cout, geom.cin):
            raise ValueError(f"conv dgrad: channels {(sc, dc)} != {(geom.cout, geom.cin)}")
    if geom.out_spatial(*in_sp) != tuple(out_sp):
        raise ValueError(f"conv: spatial extents {in_sp} -> {out_sp} inconsistent with {geom}")
    if _pad_safe(src) and _pad_safe(dst) and _pad_safe(residual):
        flags |= _lib.CONV_PADDED_CHANNELS
    d = geom.desc(n, in_sp, out_sp, x_ld, y_ld, r_ld, src.dtype, flags)
    name = ("b200seg_convtr_" if geom.transposed else "b200seg_conv_") + ("dgrad" if data_grad else "fprop")
    fn = getattr(lib, name)
    if data_grad:
        rc = fn(C.byref(d), src.data_ptr(), wp.data_ptr(), _ptr(residual), dst.data_ptr(), _stream())
    else:
        if bias is not None and (bias.dtype != torch.float32 or not bias.is_contiguous()):
            raise ValueError("conv: bias must be contiguous float32")
        rc = fn(C.byref(d), src.data_ptr(), wp.data_ptr(), _ptr(bias), _ptr(residual), dst.data_ptr(),
                _stream())
    _lib.check(rc, name)
    return dst


def conv_fprop(geom, x, wp, bias, y, residual=None, flags=0):
    """y = conv(x) [+ bias] [+ residual]  (Conv or ConvTranspose per ``geom``)."""
    return _conv_call(geom, x, wp, bias, residual, y, False, flags)


def conv_fprop_stats(geom: ConvGeom, x, wp, bias, y, eps: float = 1e-5, flags=0):
    """y = conv(x) + bias, and (mean, rstd) of y per (n, channel) for the InstanceNorm that follows,
    accumulated in the convolution's epilogue.  Falls back to a separate statistics pass when the
    layer runs on a kernel without the fusion."""
    lib = _lib.load()
    n, sd_, sh_, sw_, sc, s_ld = cl_info(x)
    n2, dd_, dh_, dw_, dc, d_ld = cl_info(y)
    if n != n2 or x.dtype != y.dtype or (sc, dc) != (geom.cin, geom.cout):
        raise ValueError("conv_fprop_stats: shape/dtype mismatch")
    if geom.out_spatial(sd_, sh_, sw_) != (dd_, dh_, dw_):
        raise ValueError("conv_fprop_stats: spatial extents inconsistent with the geometry")
    if _pad_safe(x) and _pad_safe(y):
        flags |= _lib.CONV_PADDED_CHANNELS
    d = geom.desc(n, (sd_, sh_, sw_), (dd_, dh_, dw_), s_ld, d_ld, 0, x.dtype, flags)
    c_out = (dc + 15) // 16 * 16 if _expand_pad(y) is not None else dc
    mean = torch.empty(n * c_out, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n * c_out, dtype=torch.float32, device=x.device)
    name = "b200seg_convtr_fprop_stats" if geom.transposed else "b200seg_conv_fprop_stats"
    ws = workspace(getattr(lib, name + "_workspace_bytes")(C.byref(d)), x.device)
    rc = getattr(lib, name)(C.byref(d), x.data_ptr(), wp.data_ptr(), _ptr(bias), y.data_ptr(), mean.data_ptr(),
                            rstd.data_ptr(), c_out, eps, ws.data_ptr(), ws.numel(), _stream())
    if rc == 1:  # B200SEG_STATS_NOT_FUSED
        return instnorm_stats(y, eps)
    _lib.check(rc, name)
    return mean, rstd


def col_geom(geom: ConvGeom) -> Optional[ConvGeom]:
    """The 1x1x1 geometry a small-Cin conv layer becomes on its im2col buffer (None: not applicable)."""
    j = geom.cin * geom.kernel ** geom.dims
    if geom.transposed or geom.cin > 4 or geom.kernel == 1 or not (8 <= j <= 32):
        return None
    return ConvGeom(geom.dims, j, geom.cout, 1, 1, False)


def im2col(geom: ConvGeom, x: torch.Tensor) -> torch.Tensor:
    """(N, *out_spatial, taps*cin) zero-padded channels-last im2col buffer of a small-Cin layer."""
    lib = _lib.load()
    n, d, h, w, c, x_ld = cl_info(x)
    if c != geom.cin:
        raise ValueError("im2col: channel mismatch")
    out_sp = geom.out_spatial(d, h, w)
    j = geom.cin * geom.kernel ** geom.dims
    col = alloc_activation(n, out_sp, j, x.dtype, x.device)
    col_ld = cl_info(col)[5]
    desc = geom.desc(n, (d, h, w), out_sp, x_ld, geom.cout, 0, x.dtype)
    _lib.check(lib.b200seg_im2col(C.byref(desc), x.data_ptr(), col.data_ptr(), col_ld, _stream()), "b200seg_im2col")
    return col


def conv_fprop_partials(geom: ConvGeom, x, wp, bias, y, flags=0):
    """y = conv(x) + bias with the InstanceNorm statistics left as per-CTA PARTIALS for
    ``instnorm_prelu_fwd_partials`` (which finalises them itself: one launch less on the forward chain).
    Returns (partials, ncls, tiles, cstat), or None when the layer ran on a kernel without the fusion."""
    lib = _lib.load()
    n, sd_, sh_, sw_, sc, s_ld = cl_info(x)
    n2, dd_, dh_, dw_, dc, d_ld = cl_info(y)
    if n != n2 or x.dtype != y.dtype or (sc, dc) != (geom.cin, geom.cout):
        raise ValueError("conv_fprop_partials: shape/dtype mismatch")
    if geom.out_spatial(sd_, sh_, sw_) != (dd_, dh_, dw_):
        raise ValueError("conv_fprop_partials: spatial extents inconsistent with the geometry")
    if _pad_safe(x) and _pad_safe(y):
        flags |= _lib.CONV_PADDED_CHANNELS
    d = geom.desc(n, (sd_, sh_, sw_), (dd_, dh_, dw_), s_ld, d_ld, 0, x.dtype, flags)
    name = "b200seg_convtr_fprop_stats" if geom.transposed else "b200seg_conv_fprop_stats"
    nbytes = getattr(lib, name + "_workspace_bytes")(C.byref(d))
    part = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=x.device)
    ncls, tiles = C.c_int32(0), C.c_int64(0)
    rc = lib.b200seg_conv_fprop_partials(C.byref(d), int(geom.transposed), x.data_ptr(), wp.data_ptr(), _ptr(bias),
                                         y.data_ptr(), part.data_ptr(), part.numel() * 4, C.byref(ncls),
                                         C.byref(tiles), _stream())
    if rc == 1:  # B200SEG_STATS_NOT_FUSED
        return None
    _lib.check(rc, "b200seg_conv_fprop_partials")
    return part, ncls.value, tiles.value, dc


def instnorm_prelu_fwd_partials(x, handle, alpha, y, residual=None, eps: float = 1e-5):
    """InstanceNorm + PReLU (+ residual) of the convolution output ``x`` from its partial statistics;
    returns (mean, rstd) for the backward pass (taken over the zero-padded channel count where ``x``,
    ``y`` and ``residual`` are all padded buffers, like ``conv_fprop_stats``)."""
    lib = _lib.load()
    part, ncls, tiles, cstat = handle
    ex = _expand_all(x, y, residual)
    if ex is not None:
        x, y, residual = ex
    y_ld = cl_info(y)[5]
    r_ld = cl_info(residual)[5] if residual is not None else 0
    if y.shape != x.shape or (residual is not None and residual.shape != x.shape):
        raise ValueError("instnorm_prelu_fwd_partials: shape mismatch")
    d, (n, c) = _norm_desc(x, y_ld, r_ld, eps)
    mean = torch.empty(n * c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n * c, dtype=torch.float32, device=x.device)
    _lib.check(lib.b200seg_instnorm_prelu_fwd_partials(C.byref(d), x.data_ptr(), part.data_ptr(), ncls, tiles, cstat,
                                                       mean.data_ptr(), rstd.data_ptr(), alpha.data_ptr(),
                                                       _ptr(residual), y.data_ptr(), _stream()),
               "b200seg_instnorm_prelu_fwd_partials")
    return mean, rstd


def conv_dgrad(geom, dy, wp, dx, residual=None, accumulate=False, flags=0):
    """dx = conv^T(dy) [+ residual] [+ dx]."""
    if accumulate:
        flags |= _lib.CONV_ACCUMULATE
    return _conv_call(geom, dy, wp, None, residual, dx, True, flags)


def conv_dgrad_instnorm_partials(geom, dy, wp, dx, norm_x, mean, rstd, alpha, residual=None, flags=0):
    """dx = dgrad(dy) [+ residual] AND the per-CTA partial sums of the InstanceNorm + PReLU backward that consumes dx
    as its output gradient (``norm_x`` = that layer's pre-norm tensor, ``mean`` / ``rstd`` / ``alpha`` its statistics
    and slope).  Returns a handle for ``instnorm_prelu_bwd_from_partials`` -- or None WITHOUT having launched anything
    where the fused kernel does not apply (the caller then runs ``conv_dgrad`` + ``instnorm_prelu_bwd``)."""
    lib = _lib.load()
    if geom.transposed or norm_x.shape != dx.shape:
        return None
    n, sd_, sh_, sw_, sc, s_ld = cl_info(dy)
    n2, dd_, dh_, dw_, dc, d_ld = cl_info(dx)
    if n != n2 or (sc, dc) != (geom.cout, geom.cin) or dy.dtype != torch.bfloat16 or dx.dtype != dy.dtype:
        return None
    r_ld = 0
    if residual is not None:
        if residual.shape != dx.shape or residual.dtype != dx.dtype:
            return None
        r_ld = cl_info(residual)[5]
    if not (_pad_safe(dy) and _pad_safe(dx) and _pad_safe(residual) and _pad_safe(norm_x)):
        return None
    stat_ld = mean.numel() // n
    if mean.numel() != n * stat_ld or rstd.numel() != mean.numel() or stat_ld < 16:
        return None
    flags |= _lib.CONV_PADDED_CHANNELS
    d = geom.desc(n, (dd_, dh_, dw_), (sd_, sh_, sw_), d_ld, s_ld, r_ld, dy.dtype, flags)
    nbytes = lib.b200seg_conv_dgrad_instnorm_partials_bytes(C.byref(d))
    if nbytes == 0:
        return None
    part = torch.empty(nbytes // 4, dtype=torch.float32, device=dy.device)
    rows = C.c_int64(0)
    rc = lib.b200seg_conv_dgrad_instnorm_partials(C.byref(d), dy.data_ptr(), wp.data_ptr(), _ptr(residual), dx.data_ptr(),
                                                  norm_x.data_ptr(), cl_info(norm_x)[5], mean.data_ptr(),
                                                  rstd.data_ptr(), stat_ld, alpha.data_ptr(), part.data_ptr(),
                                                  part.numel() * 4, C.byref(rows), _stream())
    if rc == 1:  # B200SEG_STATS_NOT_FUSED: nothing ran
        return None
    _lib.check(rc, "b200seg_conv_dgrad_instnorm_partials")
    return part, rows.value


def instnorm_prelu_bwd_from_partials(x, mean, rstd, alpha, dy, dx, handle, eps: float = 1e-5, out_dalpha=None):
    """The rest of the InstanceNorm + PReLU backward (final reduction + apply) after
    ``conv_dgrad_instnorm_partials``; same contract as ``instnorm_prelu_bwd``."""
    lib = _lib.load()
    part, rows = handle
    if dy.shape != x.shape or dx.shape != x.shape:
        raise ValueError("instnorm_prelu_bwd_from_partials: shape mismatch")
    mean, rstd, x, dy, dx = _match_stats(mean, rstd, "instnorm_prelu_bwd_from_partials", x, dy, dx)
    d, _ = _norm_desc(x, cl_info(dy)[5], cl_info(dx)[5], eps)
    dalpha = (torch.empty(1, dtype=torch.float32, device=x.device) if out_dalpha is None
              else _grad_out(out_dalpha, (1,), "instnorm_prelu_bwd dalpha"))
    ws = workspace(lib.b200seg_instnorm_workspace_bytes(C.byref(d)), x.device)
    _lib.check(lib.b200seg_instnorm_prelu_bwd_from_partials(C.byref(d), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                                            alpha.data_ptr(), dy.data_ptr(), part.data_ptr(), rows,
                                                            dx.data_ptr(), dalpha.data_ptr(), ws.data_ptr(), ws.numel(),
                                                            _stream()),
               "b200seg_instnorm_prelu_bwd_from_partials")
    return dalpha


def _grad_out(out, shape, what):
    if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != math.prod(shape):
        raise ValueError(f"{what}: destination must be a contiguous fp32 tensor of {math.prod(shape)} elements")
    return out


def conv_wgrad(geom: ConvGeom, x, dy, want_bias: bool = True, flags=0, out_w=None, out_b=None):
    """(gw, gbias): fp32 gradients in PyTorch parameter layout; ``out_w`` / ``out_b`` (contiguous fp32,
    e.g. views of a flat gradient bucket) receive them in place."""
    lib = _lib.load()
    n, xd, xh, xw, xc, x_ld = cl_info(x)
    n2, yd, yh, yw, yc, y_ld = cl_info(dy)
    if n != n2 or (xc, yc) != (geom.cin, geom.cout) or x.dtype != dy.dtype:
        raise ValueError("conv wgrad: shape/dtype mismatch")
    if _pad_safe(x) and _pad_safe(dy):
        flags |= _lib.CONV_PADDED_CHANNELS
    d = geom.desc(n, (xd, xh, xw), (yd, yh, yw), x_ld, y_ld, 0, x.dtype, flags)
    k = geom.kernel
    ks = (k, k) if geom.dims == 2 else (k, k, k)
    shape = (geom.cin, geom.cout, *ks) if geom.transposed else (geom.cout, geom.cin, *ks)
    gw = torch.empty(shape, dtype=torch.float32, device=x.device) if out_w is None else _grad_out(out_w, shape, "conv wgrad")
    gb = None
    if want_bias:
        gb = (torch.empty(geom.cout, dtype=torch.float32, device=x.device) if out_b is None
              else _grad_out(out_b, (geom.cout,), "conv wgrad bias"))
    pre = "b200seg_convtr_wgrad" if geom.transposed else "b200seg_conv_wgrad"
    nbytes = getattr(lib, pre + "_workspace_bytes")(C.byref(d))
    ws = workspace(nbytes, x.device)
    rc = getattr(lib, pre)(C.byref(d), x.data_ptr(), dy.data_ptr(), gw.data_ptr(), _ptr(gb),
                           ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, pre)
    return gw, gb


# ----------------------------------------------------------------------------------------------
# InstanceNorm + PReLU
# ----------------------------------------------------------------------------------------------
def _expand_pad(t: Optional[torch.Tensor]):
    """Zero-padded buffer (see ``alloc_activation``) viewed with its padding channels included, so
    the bandwidth-bound kernels run their 16-byte vector path on e.g. the 10-class tensors.  The
    padding stays zero under InstanceNorm + PReLU (x = 0 -> mean 0, xhat 0, output 0, gradient 0)."""
    if t is None or t.shape[-1] % 16 == 0 or not _pad_safe(t):
        return None
    cp = (t.shape[-1] + 15) // 16 * 16
    return t.as_strided(tuple(t.shape[:-1]) + (cp,), t.stride())


def _expand_all(*ts):
    """All given (non-None) tensors expanded, or None if any of them cannot be."""
    out = []
    for t in ts:
        if t is None:
            out.append(None)
            continue
        e = _expand_pad(t)
        if e is None:
            return None
        out.append(e)
    return out


def _norm_desc(x, y_ld, r_ld, eps):
    n, d, h, w, c, x_ld = cl_info(x)
    return NormDesc(n, c, d * h * w, x_ld, y_ld, r_ld, dtype_code(x.dtype), eps), (n, c)


def instnorm_stats(x, eps: float = 1e-5):
    lib = _lib.load()
    ex = _expand_pad(x)
    if ex is not None:
        x = ex  # statistics for C_pad channels; fwd / bwd recognise that from mean.numel()
    d, (n, c) = _norm_desc(x, 0, 0, eps)
    mean = torch.empty(n * c, dtype=torch.float32, device=x.device)
    rstd = torch.empty(n * c, dtype=torch.float32, device=x.device)
    nbytes = lib.b200seg_instnorm_workspace_bytes(C.byref(d))
    ws = workspace(nbytes, x.device)
    _lib.check(lib.b200seg_instnorm_stats(C.byref(d), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                          ws.data_ptr(), ws.numel(), _stream()), "b200seg_instnorm_stats")
    return mean, rstd


def _match_stats(mean, rstd, what, x, *others):
    """Statistics may have been taken over the zero-padded channel count: expand every operand the
    same way, or (if one of them is a concat slice) cut the statistics back to C channels."""
    n, c = x.shape[0], x.shape[-1]
    if mean.numel() == n * c:
        return (mean, rstd, x, *others)
    ex = _expand_all(x, *others)
    if ex is not None and mean.numel() == n * ex[0].shape[-1]:
        return (mean, rstd, *ex)
    cs = mean.numel() // n  # some operand is a concat slice: run on C channels with sliced statistics
    if cs < c:
        raise ValueError(f"{what}: statistics for {cs} channels do not match the operands")
    return (mean.view(n, cs)[:, :c].contiguous(), rstd.view(n, cs)[:, :c].contiguous(), x, *others)


def instnorm_prelu_fwd(x, mean, rstd, alpha, y, residual=None, eps: float = 1e-5):
    lib = _lib.load()
    y_out = y
    mean, rstd, x, y, residual = _match_stats(mean, rstd, "instnorm_prelu_fwd", x, y, residual)
    y_ld = cl_info(y)[5]
    r_ld = cl_info(residual)[5] if residual is not None else 0
    if y.shape != x.shape or (residual is not None and residual.shape != x.shape):
        raise ValueError("instnorm_prelu_fwd: shape mismatch")
    d, _ = _norm_desc(x, y_ld, r_ld, eps)
    _lib.check(lib.b200seg_instnorm_prelu_fwd(C.byref(d), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              alpha.data_ptr(), _ptr(residual), y.data_ptr(), _stream()),
               "b200seg_instnorm_prelu_fwd")
    return y_out


def instnorm_prelu_bwd(x, mean, rstd, alpha, dy, dx, eps: float = 1e-5, out_dalpha=None):
    """Writes dx; returns dalpha (1-element fp32 tensor; ``out_dalpha`` receives it in place)."""
    lib = _lib.load()
    if dy.shape != x.shape or dx.shape != x.shape:
        raise ValueError("instnorm_prelu_bwd: shape mismatch")
    mean, rstd, x, dy, dx = _match_stats(mean, rstd, "instnorm_prelu_bwd", x, dy, dx)
    d, _ = _norm_desc(x, cl_info(dy)[5], cl_info(dx)[5], eps)
    dalpha = (torch.empty(1, dtype=torch.float32, device=x.device) if out_dalpha is None
              else _grad_out(out_dalpha, (1,), "instnorm_prelu_bwd dalpha"))
    nbytes = lib.b200seg_instnorm_workspace_bytes(C.byref(d))
    ws = workspace(nbytes, x.device)
    _lib.check(lib.b200seg_instnorm_prelu_bwd(C.byref(d), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                              alpha.data_ptr(), dy.data_ptr(), dx.data_ptr(),
                                              dalpha.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "b200seg_instnorm_prelu_bwd")
    return dalpha


# ----------------------------------------------------------------------------------------------
# softmax + Dice, label maps
# ----------------------------------------------------------------------------------------------
def _labels(labels: torch.Tensor, n: int, spatial: int):
    if labels.dtype == torch.uint8:
        code = _lib.LABEL_U8
    else:
        if labels.dtype != torch.int64:
            labels = labels.long()  # MONAI does target.long()
        code = _lib.LABEL_I64
    if not labels.is_contiguous():
        labels = labels.contiguous()
    if labels.numel() != n * spatial:
        raise ValueError(f"labels have {labels.numel()} elements, expected {n * spatial}")
    return labels, code


def _dice_desc(logits, label_code, include_background=False):
    n, d, h, w, c, ld = cl_info(logits)
    return DiceDesc(n, c, d * h * w, ld, dtype_code(logits.dtype), label_code, int(include_background))


def softmax_dice_sums(logits, labels) -> torch.Tensor:
    """(N, C, 3) fp32: I = sum p*t, G = sum t, P = sum p per sample and class."""
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    desc = _dice_desc(logits, code)
    sums = torch.empty(n, c, 3, dtype=torch.float32, device=logits.device)
    nbytes = lib.b200seg_softmax_dice_workspace_bytes(C.byref(desc))
    ws = workspace(nbytes, logits.device)
    _lib.check(lib.b200seg_softmax_dice_fwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(),
                                            sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "b200seg_softmax_dice_fwd")
    return sums


def softmax_dice_metric_sums(logits, labels):
    """((N, C, 3) fp32 Dice sums, (N, C, 3) int64 metric counts {tp, |pred|, |target|}) from ONE pass over the logits."""
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    desc = _dice_desc(logits, code)
    sums = torch.empty(n, c, 3, dtype=torch.float32, device=logits.device)
    counts = torch.empty(n, c, 3, dtype=torch.int64, device=logits.device)
    ws = workspace(lib.b200seg_softmax_dice_metric_workspace_bytes(C.byref(desc)), logits.device)
    _lib.check(lib.b200seg_softmax_dice_metric_fwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(), sums.data_ptr(),
                                                   counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "b200seg_softmax_dice_metric_fwd")
    return sums, counts


def softmax_loss_sums(logits, labels, gamma: float = 2.0) -> torch.Tensor:
    """(N, C, 5) fp32: I, G, P (Dice), F = sum t (1-p)^gamma (-log p) (Focal), N = sum t (-log p) (CE)."""
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    desc = _dice_desc(logits, code)
    sums = torch.empty(n, c, 5, dtype=torch.float32, device=logits.device)
    ws = workspace(lib.b200seg_softmax_loss_workspace_bytes(C.byref(desc)), logits.device)
    _lib.check(lib.b200seg_softmax_loss_fwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(), float(gamma),
                                            sums.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
               "b200seg_softmax_loss_fwd")
    return sums


def softmax_loss_bwd(logits, labels, gamma, g_i, g_p, g_f, g_n) -> torch.Tensor:
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    if ld > c:  # keep the (zero) channel padding of the logits buffer
        dlogits = alloc_activation(n, (d, h, w), c, logits.dtype, logits.device)
        if cl_info(dlogits)[5] != ld:
            dlogits = torch.zeros((n, d, h, w, ld), dtype=logits.dtype, device=logits.device)[..., :c]
    else:
        dlogits = torch.empty(logits.shape, dtype=logits.dtype, device=logits.device)
    gs = [g.contiguous().float() for g in (g_i, g_p, g_f, g_n)]
    desc = _dice_desc(logits, code)
    _lib.check(lib.b200seg_softmax_loss_bwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(), float(gamma),
                                            gs[0].data_ptr(), gs[1].data_ptr(), gs[2].data_ptr(), gs[3].data_ptr(),
                                            dlogits.data_ptr(), _stream()), "b200seg_softmax_loss_bwd")
    return dlogits


def _dist_maps(dist: torch.Tensor, n: int, c: int, spatial: int) -> torch.Tensor:
    """Planar (N, C-1, spatial) fp32 distance maps as the dataset delivers them ((N, C-1, *S), any float dtype)."""
    if not dist.is_cuda:
        raise RuntimeError("b200seg ops run on CUDA tensors only (no CPU fallback)")
    if dist.shape[0] != n or dist.shape[1] != c - 1 or dist[0, 0].numel() != spatial:
        raise ValueError(f"distance maps {tuple(dist.shape)} do not match logits (N={n}, C-1={c - 1}, {spatial} voxels)")
    return dist.reshape(n, c - 1, spatial).float().contiguous()


def softmax_boundary_loss_sums(logits, labels, dist, gamma: float = 2.0) -> torch.Tensor:
    """(N, C, 6) fp32: softmax_loss_sums plus B = sum_v p_c * dist_{c-1} (Boundary loss; B[:, 0] = 0)."""
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    dist = _dist_maps(dist, n, c, d * h * w)
    desc = _dice_desc(logits, code)
    sums = torch.empty(n, c, 6, dtype=torch.float32, device=logits.device)
    ws = workspace(lib.b200seg_softmax_boundary_loss_workspace_bytes(C.byref(desc)), logits.device)
    _lib.check(lib.b200seg_softmax_boundary_loss_fwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(),
                                                     dist.data_ptr(), float(gamma), sums.data_ptr(), ws.data_ptr(),
                                                     ws.numel(), _stream()), "b200seg_softmax_boundary_loss_fwd")
    return sums


def softmax_boundary_loss_bwd(logits, labels, dist, gamma, g_i, g_p, g_f, g_n, g_b) -> torch.Tensor:
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    dist = _dist_maps(dist, n, c, d * h * w)
    if ld > c:  # keep the (zero) channel padding of the logits buffer
        dlogits = alloc_activation(n, (d, h, w), c, logits.dtype, logits.device)
        if cl_info(dlogits)[5] != ld:
            dlogits = torch.zeros((n, d, h, w, ld), dtype=logits.dtype, device=logits.device)[..., :c]
    else:
        dlogits = torch.empty(logits.shape, dtype=logits.dtype, device=logits.device)
    gs = [g.contiguous().float() for g in (g_i, g_p, g_f, g_n, g_b)]
    desc = _dice_desc(logits, code)
    _lib.check(lib.b200seg_softmax_boundary_loss_bwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(),
                                                     dist.data_ptr(), float(gamma), gs[0].data_ptr(),
                                                     gs[1].data_ptr(), gs[2].data_ptr(), gs[3].data_ptr(),
                                                     gs[4].data_ptr(), dlogits.data_ptr(), _stream()),
               "b200seg_softmax_boundary_loss_bwd")
    return dlogits


def dice_loss_epilogue(sums: torch.Tensor, include_background: bool, smooth: float, mean: bool):
    """(loss scalar, gI (N, C), gP (N, C)) from the (N, C, 3) Dice sums -- one launch."""
    lib = _lib.load()
    n, c, _ = sums.shape
    loss = torch.empty((), dtype=torch.float32, device=sums.device)
    g_i = torch.empty(n, c, dtype=torch.float32, device=sums.device)
    g_p = torch.empty(n, c, dtype=torch.float32, device=sums.device)
    _lib.check(lib.b200seg_dice_loss_epilogue(sums.data_ptr(), n, c, int(include_background), float(smooth),
                                              int(mean), loss.data_ptr(), g_i.data_ptr(), g_p.data_ptr(), _stream()),
               "b200seg_dice_loss_epilogue")
    return loss, g_i, g_p


def softmax_dice_bwd(logits, labels, g_i, g_p, dlogits=None) -> torch.Tensor:
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    labels, code = _labels(labels, n, d * h * w)
    if dlogits is None:
        if ld > c:  # keep the (zero) channel padding of the logits buffer
            dlogits = alloc_activation(n, (d, h, w), c, logits.dtype, logits.device)
            if cl_info(dlogits)[5] != ld:
                dlogits = torch.zeros((n, d, h, w, ld), dtype=logits.dtype, device=logits.device)[..., :c]
        else:
            dlogits = torch.empty(logits.shape, dtype=logits.dtype, device=logits.device)
    if cl_info(dlogits)[5] != ld:
        raise ValueError("softmax_dice_bwd: dlogits must share the logits' voxel stride")
    g_i = g_i.contiguous().float()
    g_p = g_p.contiguous().float()
    desc = _dice_desc(logits, code)
    _lib.check(lib.b200seg_softmax_dice_bwd(C.byref(desc), logits.data_ptr(), labels.data_ptr(),
                                            g_i.data_ptr(), g_p.data_ptr(), dlogits.data_ptr(), _stream()),
               "b200seg_softmax_dice_bwd")
    return dlogits


def argmax_dice_counts(logits, target=None, want_pred=True):
    """(pred uint8 (N, D, H, W) or None, counts int64 (N, C, 3) or None)."""
    lib = _lib.load()
    n, d, h, w, c, ld = cl_info(logits)
    code, tptr, counts = _lib.LABEL_U8, None, None
    if target is not None:
        target, code = _labels(target, n, d * h * w)
        tptr = target.data_ptr()
        counts = torch.empty(n, c, 3, dtype=torch.int64, device=logits.device)
    pred = torch.empty(n, d, h, w, dtype=torch.uint8, device=logits.device) if want_pred else None
    desc = _dice_desc(logits, code)
    _lib.check(lib.b200seg_argmax_dice_counts(C.byref(desc), logits.data_ptr(), tptr, _ptr(pred),
                                              _ptr(counts), _stream()), "b200seg_argmax_dice_counts")
    return pred, counts


def label_dice_counts(pred, target, n_classes: int) -> torch.Tensor:
    lib = _lib.load()
    n = pred.shape[0]
    spatial = pred.numel() // n
    if pred.dtype != torch.uint8:
        pred = pred.to(torch.uint8)
    pred = pred.contiguous()
    target, code = _labels(target, n, spatial)
    counts = torch.empty(n, n_classes, 3, dtype=torch.int64, device=pred.device)
    _lib.check(lib.b200seg_label_dice_counts(n, spatial, n_classes, pred.data_ptr(), target.data_ptr(),
                                             code, counts.data_ptr(), _stream()), "b200seg_label_dice_counts")
    return counts


def squash_masks(masks: torch.Tensor) -> torch.Tensor:
    """(N, S, *spatial) uint8 binary masks -> (N, *spatial) uint8 label map."""
    lib = _lib.load()
    if masks.dtype != torch.uint8:
        masks = masks.to(torch.uint8)
    masks = masks.contiguous()
    n, s = masks.shape[:2]
    spatial = masks[0, 0].numel()
    out = torch.empty((n,) + tuple(masks.shape[2:]), dtype=torch.uint8, device=masks.device)
    _lib.check(lib.b200seg_squash_masks(n, s, spatial, masks.data_ptr(), out.data_ptr(), _stream()),
               "b200seg_squash_masks")
    return out


def hu_window_norm(hu: torch.Tensor, lo, hi, mean, std, dtype=torch.float32) -> torch.Tensor:
    """int16 HU (any shape) -> (*shape, n_windows) windowed + normalised channels-last tensor."""
    lib = _lib.load()
    if hu.dtype != torch.int16:
        raise TypeError("hu_window_norm takes int16 Hounsfield units")
    hu = hu.contiguous()
    k = len(lo)
    out = torch.empty(tuple(hu.shape) + (k,), dtype=dtype, device=hu.device)
    arr = lambda v: (C.c_float * k)(*[float(a) for a in v])
    _lib.check(lib.b200seg_hu_window_norm(hu.numel(), k, hu.data_ptr(), arr(lo), arr(hi), arr(mean),
                                          arr(std), out.data_ptr(), k, dtype_code(dtype), _stream()),
               "b200seg_hu_window_norm")
    return out
