"""TEST-ONLY emulation of ``ct_image_segmentation_b200.ops`` with torch CPU arithmetic.

Used by ``tests/test_engine_plan_cpu.py`` to check the HOST logic of the UNet engine (forward and
reverse plans, zero-copy concatenation, in-place gradient fan-in) without a GPU.  Never imported
by product code.
"""
import torch
import torch.nn.functional as F

from ct_image_segmentation_b200.ops import ConvGeom, from_channels_last  # noqa: F401


def alloc_activation(n, spatial, c, dtype, device):
    return torch.empty((n, *spatial, c), dtype=dtype, device=device)


def alloc_like(t):
    return torch.empty(t.shape, dtype=t.dtype, device=t.device)


def to_channels_last(x, dtype=None):
    if x.dim() == 4:
        x = x.unsqueeze(2)
    y = x.permute(0, 2, 3, 4, 1)
    if dtype is not None:
        y = y.to(dtype)
    return y.contiguous()


def _nc(t, dims):  # (N,D,H,W,C) -> (N,C,D,H,W) or (N,C,H,W)
    t = t.permute(0, 4, 1, 2, 3)
    return t.squeeze(2) if dims == 2 else t


def _cl(t, dims):
    if dims == 2:
        t = t.unsqueeze(2)
    return t.permute(0, 2, 3, 4, 1)


def pack_weight(geom, kind, w, dtype):
    return w.detach().to(dtype).to(torch.float32)  # weights are stored in the compute dtype


def _conv(geom, x, w, bias=None):
    p = (geom.kernel - 1) // 2
    if geom.transposed:
        f = F.conv_transpose2d if geom.dims == 2 else F.conv_transpose3d
        return f(x, w, bias, stride=geom.stride, padding=p, output_padding=geom.stride - 1)
    f = F.conv2d if geom.dims == 2 else F.conv3d
    return f(x, w, bias, stride=geom.stride, padding=p)


def conv_fprop(geom, x, wp, bias, y, residual=None, flags=0):
    out = _cl(_conv(geom, _nc(x, geom.dims).float(), wp, bias), geom.dims)
    if residual is not None:
        out = out + residual.float()
    y.copy_(out)
    return y


def conv_fprop_stats(geom, x, wp, bias, y, eps=1e-5, flags=0):
    conv_fprop(geom, x, wp, bias, y)
    return instnorm_stats(y, eps)


def conv_dgrad(geom, dy, wp, dx, residual=None, accumulate=False, flags=0):
    xin = torch.zeros(_nc(dx, geom.dims).shape, dtype=torch.float32, requires_grad=True)
    out = _conv(geom, xin, wp)
    (g,) = torch.autograd.grad(out, xin, _nc(dy, geom.dims).float())
    g = _cl(g, geom.dims)
    if residual is not None:
        g = g + residual.float()
    if accumulate:
        g = g + dx.float()
    dx.copy_(g)
    return dx


def conv_wgrad(geom, x, dy, want_bias=True, flags=0, out_w=None, out_b=None):
    k = geom.kernel
    ks = (k, k) if geom.dims == 2 else (k, k, k)
    shape = (geom.cin, geom.cout, *ks) if geom.transposed else (geom.cout, geom.cin, *ks)
    w = torch.zeros(shape, requires_grad=True)
    out = _conv(geom, _nc(x, geom.dims).float(), w)
    (gw,) = torch.autograd.grad(out, w, _nc(dy, geom.dims).float())
    gb = dy.float().sum(dim=(0, 1, 2, 3)) if want_bias else None
    if out_w is not None:
        out_w.view(-1).copy_(gw.reshape(-1))
        gw = out_w
    if out_b is not None and gb is not None:
        out_b.copy_(gb)
        gb = out_b
    return gw, gb


def instnorm_stats(x, eps=1e-5):
    xf = x.float()
    mean = xf.mean(dim=(1, 2, 3))
    var = xf.var(dim=(1, 2, 3), unbiased=False)
    return mean.reshape(-1), (1.0 / torch.sqrt(var + eps)).reshape(-1)


def instnorm_prelu_fwd(x, mean, rstd, alpha, y, residual=None, eps=1e-5):
    n, c = x.shape[0], x.shape[4]
    h = (x.float() - mean.view(n, 1, 1, 1, c)) * rstd.view(n, 1, 1, 1, c)
    out = torch.where(h > 0, h, alpha.float() * h)
    if residual is not None:
        out = out + residual.float()
    y.copy_(out)
    return y


def instnorm_prelu_bwd(x, mean, rstd, alpha, dy, dx, eps=1e-5, out_dalpha=None):
    n, c = x.shape[0], x.shape[4]
    r = rstd.view(n, 1, 1, 1, c)
    h = (x.float() - mean.view(n, 1, 1, 1, c)) * r
    g = dy.float()
    pos = h > 0
    gt = torch.where(pos, g, alpha.float() * g)
    dalpha = torch.where(pos, torch.zeros_like(g), g * h).sum().reshape(1)
    s1 = gt.mean(dim=(1, 2, 3), keepdim=True)
    s2 = (gt * h).mean(dim=(1, 2, 3), keepdim=True)
    dx.copy_(r * (gt - s1 - h * s2))
    if out_dalpha is not None:
        out_dalpha.copy_(dalpha)
        return out_dalpha
    return dalpha


def conv_dgrad_instnorm_partials(*args, **kwargs):
    return None  # (the emulation has no fused kernel: the plan falls back to conv_dgrad + instnorm_prelu_bwd)
