"""Per-kernel parity on the GPU, through the C ABI (ctypes) -- `pytest -m gpu`.

Oracle: torch CPU fp32 ops (convolutions, InstanceNorm, PReLU) and oracle/monai_ref.py; golden
fixtures produced by the reference's own in-tree functions (tests/golden).  Tolerances per
BASELINE.json north_star: 1e-4 relative (fp32 check mode), 1e-2 relative (bf16); integer / label
outputs bit-exact.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ct_image_segmentation_b200 import _lib, losses, metrics, ops, transforms
from ct_image_segmentation_b200.ops import ConvGeom
from oracle import monai_ref as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = {torch.float32: 1e-4, torch.bfloat16: 1e-2}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def q(t, dtype):
    """Round a CPU fp32 tensor to the storage dtype (so both sides see identical inputs)."""
    return t.to(dtype).float()


def cl_dev(t_nc, dtype, pad_c=0, c_off=0):
    """CPU (N,C,*S) fp32 -> CUDA channels-last (N,D,H,W,C) tensor; with pad_c > 0 the result is a
    channel slice [c_off, c_off+C) of a wider buffer (exercises ld > C)."""
    if t_nc.dim() == 4:
        t_nc = t_nc.unsqueeze(2)
    cl = t_nc.permute(0, 2, 3, 4, 1).contiguous().to(DEV, dtype)
    if pad_c == 0:
        return cl
    buf = torch.full(cl.shape[:-1] + (cl.shape[-1] + pad_c,), 7.0, dtype=dtype, device=DEV)
    view = buf[..., c_off:c_off + cl.shape[-1]]
    view.copy_(cl)
    return view


def nc_cpu(t_cl, dims):
    t = t_cl.float().cpu().permute(0, 4, 1, 2, 3)
    return t.squeeze(2) if dims == 2 else t


def ref_conv(g: ConvGeom, x, w, b=None):
    p = (g.kernel - 1) // 2
    if g.transposed:
        f = F.conv_transpose2d if g.dims == 2 else F.conv_transpose3d
        return f(x, w, b, stride=g.stride, padding=p, output_padding=g.stride - 1)
    f = F.conv2d if g.dims == 2 else F.conv3d
    return f(x, w, b, stride=g.stride, padding=p)


GEOMS = [
    # dims, cin, cout, k, stride, transposed, spatial
    (3, 1, 16, 3, 2, False, (8, 12, 16)),
    (3, 16, 16, 3, 1, False, (6, 8, 10)),
    (3, 16, 32, 3, 2, False, (8, 8, 12)),
    (3, 32, 10, 3, 2, True, (4, 6, 8)),
    (3, 10, 10, 3, 1, False, (6, 6, 10)),
    (3, 24, 40, 1, 1, False, (4, 4, 6)),
    (3, 96, 32, 3, 2, True, (3, 4, 5)),
    (3, 64, 80, 3, 1, False, (4, 4, 4)),
    (2, 3, 8, 3, 2, False, (20, 24)),
    (2, 16, 8, 3, 2, True, (10, 12)),
    (2, 8, 8, 3, 1, False, (9, 16)),
    # first-layer shapes of the three configurations (register-tiled small-Cin wgrad, ragged chunks)
    (3, 1, 16, 3, 2, False, (18, 20, 22)),
    (3, 1, 32, 3, 2, False, (10, 12, 14)),
    (2, 1, 64, 3, 2, False, (30, 36)),
    (2, 3, 16, 3, 1, False, (17, 19)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims,cin,cout,k,s,tr,sp", GEOMS)
def test_conv_fprop_dgrad_wgrad(dims, cin, cout, k, s, tr, sp, dtype):
    torch.manual_seed(12342)
    g = ConvGeom(dims, cin, cout, k, s, tr)
    n = 2
    ks = (k,) * dims
    w = q(torch.randn((cin, cout, *ks) if tr else (cout, cin, *ks)) * 0.2, dtype)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y_ref = ref_conv(g, x, w, b)
    res = q(torch.randn_like(y_ref), dtype)
    dy = q(torch.randn_like(y_ref), dtype)
    (y_ref + res).backward(dy)

    wdev = w.detach().to(DEV)
    x_cl = cl_dev(x.detach(), dtype, pad_c=8, c_off=8)           # strided source
    res_cl = cl_dev(res, dtype)
    # fp32: odd channel offset (unaligned scalar path); bf16: 16-byte aligned slice (tcgen05 path)
    off = 3 if dtype == torch.float32 else 8
    y_buf = torch.zeros(res_cl.shape[:-1] + (cout + 16,), dtype=dtype, device=DEV)
    y_cl = y_buf[..., off:off + cout]
    kind_f = _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP
    kind_d = _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD
    ops.conv_fprop(g, x_cl, ops.pack_weight(g, kind_f, wdev, dtype), b.to(DEV), y_cl, res_cl)
    tol = TOL[dtype]
    e = rel(nc_cpu(y_cl, dims), (y_ref + res).detach())
    assert e < tol, f"fprop rel err {e}"
    assert float(y_buf[..., :off].abs().max()) == 0.0  # neighbours of the slice untouched
    assert float(y_buf[..., off + cout:].abs().max()) == 0.0

    # dgrad: plain, then accumulate + residual
    dy_cl = cl_dev(dy, dtype, pad_c=8, c_off=0)
    dx_cl = torch.empty(x_cl.shape, dtype=dtype, device=DEV)
    wp_d = ops.pack_weight(g, kind_d, wdev, dtype)
    ops.conv_dgrad(g, dy_cl, wp_d, dx_cl)
    e = rel(nc_cpu(dx_cl, dims), x.grad)
    assert e < tol, f"dgrad rel err {e}"
    addend = q(torch.randn_like(x.grad), dtype)
    base = q(torch.randn_like(x.grad), dtype)
    dx2 = cl_dev(base, dtype, pad_c=8, c_off=8)
    ops.conv_dgrad(g, dy_cl, wp_d, dx2, residual=cl_dev(addend, dtype), accumulate=True)
    e = rel(nc_cpu(dx2, dims), x.grad + addend + base)
    assert e < max(tol, 2e-2 if dtype == torch.bfloat16 else 0), f"dgrad accumulate rel err {e}"

    gw, gb = ops.conv_wgrad(g, x_cl, dy_cl)
    e = rel(gw, w.grad)
    assert e < tol, f"wgrad rel err {e}"
    e = rel(gb, dy.sum(dim=[0] + list(range(2, 2 + dims))))
    assert e < tol, f"bias grad rel err {e}"


def test_pack_weights_batched_matches_single():
    """One-launch repack of many layers == per-layer b200seg_pack_weight, byte for byte."""
    torch.manual_seed(3)
    dtype = torch.bfloat16
    layers = [(ConvGeom(3, 1, 16, 3, 2, False), _lib.W_CONV_FPROP), (ConvGeom(3, 16, 16, 3, 1, False), _lib.W_CONV_DGRAD),
              (ConvGeom(3, 32, 10, 3, 2, True), _lib.W_CONVTR_FPROP), (ConvGeom(3, 32, 10, 3, 2, True), _lib.W_CONVTR_DGRAD),
              (ConvGeom(3, 128, 256, 3, 1, False), _lib.W_CONV_FPROP), (ConvGeom(3, 128, 256, 1, 1, False), _lib.W_CONV_DGRAD),
              (ConvGeom(2, 64, 128, 3, 2, False), _lib.W_CONV_FPROP), (ConvGeom(3, 384, 64, 3, 2, True), _lib.W_CONVTR_FPROP)]
    ws, bufs, entries, singles = [], [], [], []
    for g, kind in layers:
        ks = (g.kernel,) * g.dims
        w = torch.randn((g.cin, g.cout, *ks) if g.transposed else (g.cout, g.cin, *ks), device=DEV)
        nbytes, tc_off = ops.packed_weight_layout(g, kind, dtype)
        buf = torch.full((nbytes,), 0x5A, dtype=torch.uint8, device=DEV)
        ws.append(w); bufs.append(buf)
        entries.append((w.data_ptr(), buf.data_ptr(), tc_off, g.kernel ** g.dims, g.cin, g.cout, kind))
        singles.append(ops.pack_weight(g, kind, w, dtype))
    table = ops.make_pack_table(entries, torch.device(DEV))
    ops.pack_weights_batched(table, len(entries))
    torch.cuda.synchronize()
    for (g, kind), buf, single in zip(layers, bufs, singles):
        sb = single if single.dtype == torch.uint8 else single.view(torch.uint8)
        nb, tc_off = ops.packed_weight_layout(g, kind, dtype)
        taps = g.kernel ** g.dims
        gen_bytes = taps * g.cin * g.cout * 2
        assert torch.equal(buf[:gen_bytes], sb.reshape(-1)[:gen_bytes]), f"generic layout differs for {g}"
        assert torch.equal(buf[tc_off:], sb.reshape(-1)[tc_off:nb]), f"tcgen05 layout differs for {g}"


TC_GEOMS = [
    # dims, cin, cout, k, stride, transposed, n, spatial  -- shapes the tcgen05 kernels take
    (3, 16, 16, 3, 1, False, 2, (16, 16, 16)),
    (3, 16, 32, 3, 2, False, 2, (16, 24, 16)),
    (3, 32, 32, 3, 1, False, 1, (8, 12, 20)),
    (3, 64, 16, 3, 2, True, 1, (6, 8, 8)),
    (3, 32, 10, 3, 2, True, 1, (8, 8, 8)),       # 10 classes: zero-padded to 16 channels
    (3, 10, 10, 3, 1, False, 1, (16, 16, 16)),
    (3, 128, 256, 3, 1, False, 2, (6, 6, 6)),    # 2 K blocks, 2 N tiles, ragged tiles
    (3, 128, 256, 1, 1, False, 1, (8, 8, 8)),    # 1x1x1 residual conv
    (3, 384, 64, 3, 2, True, 1, (4, 4, 4)),
    (3, 256, 256, 3, 1, False, 1, (8, 8, 8)),
    (2, 64, 128, 3, 2, False, 2, (32, 48)),
    (2, 128, 64, 3, 2, True, 1, (16, 24)),
    (2, 16, 16, 3, 1, False, 1, (40, 56)),
    # sliding-window kernels (3x3x3 stride 1, 16/32 channels, larger volumes, ragged tiles, d segments)
    (3, 16, 16, 3, 1, False, 1, (12, 40, 24)),
    (3, 32, 32, 3, 1, False, 2, (8, 32, 32)),
    (3, 10, 10, 3, 1, False, 1, (20, 32, 40)),
    (3, 16, 32, 3, 1, False, 1, (9, 48, 16)),
    (3, 32, 16, 3, 1, False, 1, (16, 24, 32)),
    (3, 16, 16, 3, 1, False, 2, (40, 128, 128)),   # long sweeps: the TMEM accumulator ring wraps
    # line-tiled kernel (tc_line.cu: rows of 32 / 64 / 128 voxels, 16 padded channels): ragged line tiles, several
    # items per persistent CTA, rows 31|32 of a 128-voxel line exchanged through shared memory
    (3, 16, 16, 3, 1, False, 2, (9, 19, 32)),
    (3, 10, 10, 3, 1, False, 1, (13, 10, 64)),
    (3, 16, 10, 3, 1, False, 1, (6, 5, 128)),
    (3, 10, 16, 3, 1, False, 3, (5, 3, 128)),
]


@pytest.mark.parametrize("dims,cin,cout,k,s,tr,n,sp", TC_GEOMS)
def test_tcgen05_conv_vs_torch_and_generic(dims, cin, cout, k, s, tr, n, sp):
    """bf16 tcgen05 kernels (fprop + dgrad, with residual / accumulate) against torch fp32 on the
    same bf16-rounded inputs (1e-2) and against the CUDA-core kernel through the same ABI."""
    lib = _lib.load()
    dtype = torch.bfloat16
    torch.manual_seed(4242)
    g = ConvGeom(dims, cin, cout, k, s, tr)
    ks = (k,) * dims
    w = q(torch.randn((cin, cout, *ks) if tr else (cout, cin, *ks)) * (2.0 / (cin * k ** dims)) ** 0.5, dtype)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype).requires_grad_(True)
    w.requires_grad_(True)
    y_ref = ref_conv(g, x, w, b)
    res = q(torch.randn_like(y_ref), dtype)
    dy = q(torch.randn_like(y_ref), dtype)
    (y_ref + res).backward(dy)

    def dev(t_nc):  # padded, pad-safe channels-last device tensor
        if t_nc.dim() == 4:
            t_nc = t_nc.unsqueeze(2)
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    wdev = w.detach().to(DEV)
    kind_f = _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP
    kind_d = _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD
    x_cl, res_cl, dy_cl = dev(x.detach()), dev(res), dev(dy)
    y_cl = ops.alloc_like(res_cl)
    t0 = lib.b200seg_tc_launch_count()
    ops.conv_fprop(g, x_cl, ops.pack_weight(g, kind_f, wdev, dtype), b.to(DEV), y_cl, res_cl)
    assert lib.b200seg_tc_launch_count() == t0 + 1, "tcgen05 kernel was not used"
    e = rel(nc_cpu(y_cl, dims), (y_ref + res).detach())
    assert e < 1e-2, f"tc fprop rel err {e}"
    y_gen = ops.alloc_like(res_cl)
    ops.conv_fprop(g, x_cl, ops.pack_weight(g, kind_f, wdev, dtype), b.to(DEV), y_gen, res_cl,
                   flags=_lib.CONV_FORCE_GENERIC)
    assert lib.b200seg_tc_launch_count() == t0 + 1
    assert rel(y_cl, y_gen) < 6e-3, "tc vs generic fprop"

    wp_d = ops.pack_weight(g, kind_d, wdev, dtype)
    dx_cl = ops.alloc_like(x_cl)
    ops.conv_dgrad(g, dy_cl, wp_d, dx_cl)
    assert lib.b200seg_tc_launch_count() == t0 + 2, "tcgen05 dgrad kernel was not used"
    e = rel(nc_cpu(dx_cl, dims), x.grad)
    assert e < 1e-2, f"tc dgrad rel err {e}"
    addend = q(torch.randn_like(x.grad), dtype)
    base = q(torch.randn_like(x.grad), dtype)
    dx2 = dev(base)
    ops.conv_dgrad(g, dy_cl, wp_d, dx2, residual=dev(addend), accumulate=True)
    e = rel(nc_cpu(dx2, dims), x.grad + addend + base)
    assert e < 2e-2, f"tc dgrad accumulate rel err {e}"
    if cout % 16:  # channel padding of the destination stays zero
        full = y_cl.as_strided(y_cl.shape[:-1] + ((cout + 15) // 16 * 16,), y_cl.stride())
        assert float(full[..., cout:].abs().max()) == 0.0

    # weight gradient on tcgen05 (MN-major operands, split-K) vs torch fp32 and vs the CUDA-core kernel
    t1 = lib.b200seg_tc_launch_count()
    gw, gb = ops.conv_wgrad(g, x_cl, dy_cl)
    assert lib.b200seg_tc_launch_count() == t1 + 1, "tcgen05 wgrad kernel was not used"
    e = rel(gw, w.grad)
    assert e < 1e-2, f"tc wgrad rel err {e}"
    gw2, _ = ops.conv_wgrad(g, x_cl, dy_cl, flags=_lib.CONV_FORCE_GENERIC)
    assert rel(gw, gw2) < 1e-4, "tc vs generic wgrad"
    assert rel(gb, dy.sum(dim=[0] + list(range(2, 2 + dims)))) < 1e-2


SPLITK_CASES = [
    # cin, cout, stride, transposed, n, input spatial -- layers of the 8^3 levels (<= 74 CTAs on the streaming kernel)
    (256, 256, 1, False, 2, (8, 8, 8)),     # 8 tiles x 8 channel tiles -> clusters of 2
    (128, 128, 1, False, 2, (8, 8, 8)),     # 8 x 4 -> clusters of 4
    (64, 128, 2, False, 2, (16, 16, 16)),   # stride 2 into 8^3 (parity maps)
    (128, 256, 1, False, 1, (6, 6, 6)),     # ragged tiles
    (384, 64, 2, True, 1, (4, 4, 4)),       # ConvTranspose: 8 parity classes with 1..8 taps each
]


@pytest.mark.parametrize("cin,cout,s,tr,n,sp", SPLITK_CASES)
def test_splitk_cluster_conv(cin, cout, s, tr, n, sp):
    """Split-K cluster kernel (B200SEG_CONV_SPLIT_K) against the streaming kernel it replaces -- same MMAs in
    another summation order: 2e-3 -- and torch fp32 (1e-2); fprop with statistics partials, dgrad with
    residual + accumulate."""
    lib = _lib.load()
    dtype = torch.bfloat16
    torch.manual_seed(4711)
    g = ConvGeom(3, cin, cout, 3, s, tr)
    w = q(torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3)) * (2.0 / (cin * 27)) ** 0.5, dtype)
    w.requires_grad_(True)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype).requires_grad_(True)
    y_ref = ref_conv(g, x, w, b)
    dy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(dy)

    def dev(t_nc):
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    wdev = w.detach().to(DEV)
    x_cl, dy_cl = dev(x.detach()), dev(dy)
    wp = ops.pack_weight(g, _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP, wdev, dtype)
    y0, y1 = ops.alloc_like(dy_cl), ops.alloc_like(dy_cl)
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y0, flags=_lib.CONV_NO_SPLIT_K)
    assert lib.b200seg_last_launch() == b"tc_conv"
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y1, flags=_lib.CONV_SPLIT_K)
    assert lib.b200seg_last_launch() == b"tc_conv_splitk"
    assert rel(y1, y0) < 2e-3 and rel(nc_cpu(y1, 3), y_ref.detach()) < 1e-2
    # statistics partials through the split kernel
    c1, c0 = ops.alloc_like(dy_cl), ops.alloc_like(dy_cl)
    h1 = ops.conv_fprop_partials(g, x_cl, wp, b.to(DEV), c1, flags=_lib.CONV_SPLIT_K)
    h0 = ops.conv_fprop_partials(g, x_cl, wp, b.to(DEV), c0, flags=_lib.CONV_NO_SPLIT_K)
    assert (h1 is None) == (h0 is None)
    if h1 is not None:
        a1, a0 = ops.alloc_like(c1), ops.alloc_like(c0)
        alpha = torch.tensor([0.25], device=DEV)
        m1, r1 = ops.instnorm_prelu_fwd_partials(c1, h1, alpha, a1)
        m0, r0 = ops.instnorm_prelu_fwd_partials(c0, h0, alpha, a0)
        assert rel(m1, m0) < 1e-4 and rel(r1, r0) < 1e-4 and rel(a1, a0) < 5e-3
    # dgrad with residual + accumulate
    wp_d = ops.pack_weight(g, _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD, wdev, dtype)
    addend, base = q(torch.randn_like(x.grad), dtype), q(torch.randn_like(x.grad), dtype)
    dx0, dx1 = dev(base), dev(base)
    ops.conv_dgrad(g, dy_cl, wp_d, dx0, residual=dev(addend), accumulate=True, flags=_lib.CONV_NO_SPLIT_K)
    ops.conv_dgrad(g, dy_cl, wp_d, dx1, residual=dev(addend), accumulate=True, flags=_lib.CONV_SPLIT_K)
    assert rel(dx1, dx0) < 4e-3 and rel(nc_cpu(dx1, 3), x.grad + addend + base) < 2e-2
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y0)   # the default dispatch (no flag) is the split kernel where it applies
    assert lib.b200seg_last_launch() == b"tc_conv_splitk"


@pytest.mark.parametrize("cin,cout,n,sp,with_res", [(16, 16, 2, (12, 40, 24), False), (10, 10, 1, (9, 32, 40), True),
                                                     (16, 16, 1, (8, 64, 64), True), (10, 10, 2, (16, 32, 32), False),
                                                     (10, 10, 2, (9, 8, 128), True), (16, 16, 1, (20, 9, 64), False)])
def test_dgrad_fused_with_instnorm_backward_sums(cin, cout, n, sp, with_res):
    """b200seg_conv_dgrad_instnorm_partials + b200seg_instnorm_prelu_bwd_from_partials (the dgrad epilogue leaves the
    three per-(sample, channel) sums of the InstanceNorm + PReLU backward its result feeds) against the two separate
    entry points: the input gradient is bit-identical, the InstanceNorm-backward result and the PReLU-slope gradient
    agree to summation order, and both match torch autograd on the same rounded inputs."""
    dtype = torch.bfloat16
    torch.manual_seed(77)
    g = ConvGeom(3, cin, cout, 3, 1, False)
    w = q(torch.randn(cout, cin, 3, 3, 3) * 0.1, dtype)
    # layer L (consumer of the gradient): c_prev -> InstanceNorm -> PReLU -> x ; layer L+1: conv(x)
    c_prev = q(torch.randn(n, cin, *sp) * 1.5 + 0.3, dtype).requires_grad_(True)
    alpha = torch.tensor([0.2], requires_grad=True)
    dy = q(torch.randn(n, cout, *sp), dtype)
    addend = q(torch.randn(n, cin, *sp), dtype) if with_res else None
    xh = F.prelu(F.instance_norm(c_prev, eps=1e-5), alpha)
    y = ref_conv(g, xh, w) + (0 if addend is None else 0)
    go = torch.autograd.grad(y, xh, dy, retain_graph=True)[0]
    go_total = q(go + (addend if with_res else 0), dtype)          # what the dgrad epilogue stores (bf16)
    (gc_ref, da_ref) = torch.autograd.grad(xh, (c_prev, alpha), go_total)

    def dev(t_nc):
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    c_d, dy_d = dev(c_prev.detach()), dev(dy)
    res_d = dev(addend) if with_res else None
    wd = ops.pack_weight(g, _lib.W_CONV_DGRAD, w.to(DEV), dtype)
    mean, rstd = ops.instnorm_stats(c_d)
    a_d = alpha.detach().to(DEV)
    lib = _lib.load()
    # separate path
    dx0, gc0 = ops.alloc_like(c_d), ops.alloc_like(c_d)
    ops.conv_dgrad(g, dy_d, wd, dx0, residual=res_d)
    da0 = ops.instnorm_prelu_bwd(c_d, mean, rstd, a_d, dx0, gc0)
    # fused path
    dx1, gc1 = ops.alloc_like(c_d), ops.alloc_like(c_d)
    h = ops.conv_dgrad_instnorm_partials(g, dy_d, wd, dx1, c_d, mean, rstd, a_d, residual=res_d)
    assert h is not None and lib.b200seg_last_launch() in (b"tc_slide_conv_bwdstats", b"tc_line_conv_bwdstats")
    da1 = ops.instnorm_prelu_bwd_from_partials(c_d, mean, rstd, a_d, dx1, gc1, h)
    assert lib.b200seg_last_launch() == b"instnorm_prelu_bwd_apply"
    assert torch.equal(dx1, dx0)
    assert rel(nc_cpu(dx1, 3), go_total) < 1e-2
    assert rel(gc1, gc0) < 2e-3 and rel(nc_cpu(gc1, 3), gc_ref) < 1e-2
    assert abs(da1.item() - da0.item()) <= 1e-3 * abs(da0.item()) + 1e-4
    assert abs(da1.item() - da_ref.item()) <= 2e-2 * abs(da_ref.item()) + 1e-2
    # a layer the fused kernel does not take: nothing is launched, None comes back
    g32 = ConvGeom(3, 32, 32, 3, 1, False)
    x32 = ops.alloc_activation(1, (8, 32, 32), 32, dtype, DEV)
    m32, r32 = ops.instnorm_stats(x32)
    before = lib.b200seg_launch_count()
    assert ops.conv_dgrad_instnorm_partials(g32, x32, ops.pack_weight(g32, _lib.W_CONV_DGRAD, torch.zeros(32, 32, 3, 3, 3, device=DEV), dtype),
                                            ops.alloc_like(x32), x32, m32, r32, a_d) is None
    assert lib.b200seg_launch_count() == before + 1   # (only the weight pack)


CONVTR_SLIDE = [
    # cin, cout, n, input spatial -- ConvTranspose k3 s2 layers the sliding-window kernels take
    (32, 10, 1, (8, 24, 40)),     # top layer: 10 classes padded to 16, ragged h tile
    (64, 16, 2, (9, 32, 16)),     # odd depth, d segments
    (32, 16, 1, (5, 16, 36)),     # ragged w tile
]


@pytest.mark.parametrize("cin,cout,n,sp", CONVTR_SLIDE)
def test_convtr_sliding_kernels(cin, cout, n, sp):
    """Sliding-window ConvTranspose kernels (fprop with fused InstanceNorm statistics, dgrad, wgrad)
    against torch fp32 on the same bf16-rounded inputs (1e-2) and against the streaming kernels."""
    lib = _lib.load()
    dtype = torch.bfloat16
    torch.manual_seed(777)
    g = ConvGeom(3, cin, cout, 3, 2, True)
    w = q(torch.randn(cin, cout, 3, 3, 3) * (2.0 / (cin * 27)) ** 0.5, dtype).requires_grad_(True)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype).requires_grad_(True)
    y_ref = ref_conv(g, x, w, b)
    dy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(dy)

    def dev(t_nc):
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    wdev = w.detach().to(DEV)
    x_cl, dy_cl = dev(x.detach()), dev(dy)
    wp = ops.pack_weight(g, _lib.W_CONVTR_FPROP, wdev, dtype)
    y_cl = ops.alloc_like(dy_cl)
    t0 = lib.b200seg_tc_launch_count()
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y_cl)
    assert lib.b200seg_tc_launch_count() == t0 + 1
    assert lib.b200seg_last_launch() == b"tc_convtr_fprop"
    e = rel(nc_cpu(y_cl, 3), y_ref.detach())
    assert e < 1e-2, f"convtr fprop rel err {e}"
    y_str = ops.alloc_like(dy_cl)
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y_str, flags=_lib.CONV_NO_SLIDE)
    assert rel(y_cl, y_str) < 2e-3, "sliding vs streaming fprop"
    if cout % 16:
        full = y_cl.as_strided(y_cl.shape[:-1] + ((cout + 15) // 16 * 16,), y_cl.stride())
        assert float(full[..., cout:].abs().max()) == 0.0
    # fused statistics
    y2 = ops.alloc_like(dy_cl)
    mean, rstd = ops.conv_fprop_stats(g, x_cl, wp, b.to(DEV), y2)
    assert torch.equal(y2, y_cl)
    yd = y_ref.detach().double()
    m_ref = yd.mean(dim=(2, 3, 4))
    r_ref = 1.0 / torch.sqrt(yd.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    cs = mean.numel() // n
    assert rel(mean.view(n, cs)[:, :cout], m_ref) < 2e-3 and rel(rstd.view(n, cs)[:, :cout], r_ref) < 2e-3
    # dgrad
    wp_d = ops.pack_weight(g, _lib.W_CONVTR_DGRAD, wdev, dtype)
    dx_cl = ops.alloc_like(x_cl)
    ops.conv_dgrad(g, dy_cl, wp_d, dx_cl)
    assert lib.b200seg_last_launch() == b"tc_convtr_dgrad"
    e = rel(nc_cpu(dx_cl, 3), x.grad)
    assert e < 1e-2, f"convtr dgrad rel err {e}"
    dx_str = ops.alloc_like(x_cl)
    ops.conv_dgrad(g, dy_cl, wp_d, dx_str, flags=_lib.CONV_NO_SLIDE)
    assert rel(dx_cl, dx_str) < 2e-3, "sliding vs streaming dgrad"
    # wgrad
    gw, gb = ops.conv_wgrad(g, x_cl, dy_cl, want_bias=False)
    if cin == 32:
        assert lib.b200seg_last_launch() == b"tc_slide_wgrad_unpack"
    e = rel(gw, w.grad)
    assert e < 1e-2, f"convtr wgrad rel err {e}"
    gw2, _ = ops.conv_wgrad(g, x_cl, dy_cl, flags=_lib.CONV_NO_SLIDE)
    assert rel(gw, gw2) < 1e-4, "sliding vs streaming wgrad"


CONV_S2_SLIDE = [
    # cin, cout, n, input spatial -- Conv k3 s2 layers the same hi/lo sliding kernels take
    (16, 32, 1, (16, 48, 64)),
    (16, 32, 2, (10, 64, 48)),
    (10, 32, 1, (8, 64, 32)),     # padded hi side
]


@pytest.mark.parametrize("cin,cout,n,sp", CONV_S2_SLIDE)
def test_conv_stride2_sliding_kernels(cin, cout, n, sp):
    """Stride-2 Conv on the hi/lo sliding kernels: fprop (+ fused statistics) = hi->lo, dgrad (plain and
    accumulating into a channel slice of a wider buffer) = lo->hi, wgrad; vs torch fp32 and streaming."""
    lib = _lib.load()
    dtype = torch.bfloat16
    torch.manual_seed(778)
    g = ConvGeom(3, cin, cout, 3, 2, False)
    w = q(torch.randn(cout, cin, 3, 3, 3) * (2.0 / (cin * 27)) ** 0.5, dtype).requires_grad_(True)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype).requires_grad_(True)
    y_ref = ref_conv(g, x, w, b)
    dy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(dy)

    def dev(t_nc):
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    wdev = w.detach().to(DEV)
    x_cl, dy_cl = dev(x.detach()), dev(dy)
    wp = ops.pack_weight(g, _lib.W_CONV_FPROP, wdev, dtype)
    y_cl = ops.alloc_like(dy_cl)
    ops.conv_fprop(g, x_cl, wp, b.to(DEV), y_cl)
    assert lib.b200seg_last_launch() == b"tc_convtr_dgrad"
    e = rel(nc_cpu(y_cl, 3), y_ref.detach())
    assert e < 1e-2, f"conv s2 fprop rel err {e}"
    y2 = ops.alloc_like(dy_cl)
    mean, rstd = ops.conv_fprop_stats(g, x_cl, wp, b.to(DEV), y2)
    assert torch.equal(y2, y_cl)
    yd = y_ref.detach().double()
    m_ref = yd.mean(dim=(2, 3, 4))
    r_ref = 1.0 / torch.sqrt(yd.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    assert rel(mean.view(n, -1)[:, :cout], m_ref) < 2e-3 and rel(rstd.view(n, -1)[:, :cout], r_ref) < 2e-3
    # dgrad, plain
    wp_d = ops.pack_weight(g, _lib.W_CONV_DGRAD, wdev, dtype)
    dx_cl = ops.alloc_like(x_cl)
    ops.conv_dgrad(g, dy_cl, wp_d, dx_cl)
    assert lib.b200seg_last_launch() == b"tc_convtr_fprop"
    e = rel(nc_cpu(dx_cl, 3), x.grad)
    assert e < 1e-2, f"conv s2 dgrad rel err {e}"
    if cin % 16 == 0:
        # accumulate into the first channels of a wider buffer (the skip half of a gradient concat buffer)
        base = q(torch.randn(n, 2 * cin, *sp), dtype)
        wide = dev(base)
        ops.conv_dgrad(g, dy_cl, wp_d, wide[..., :cin], accumulate=True)
        assert lib.b200seg_last_launch() == b"tc_convtr_fprop"
        got = nc_cpu(wide, 3)
        assert rel(got[:, :cin], x.grad + base[:, :cin]) < 2e-2
        assert torch.equal(got[:, cin:], base[:, cin:])
    # wgrad
    gw, _ = ops.conv_wgrad(g, x_cl, dy_cl, want_bias=False)
    assert lib.b200seg_last_launch() == b"tc_slide_wgrad_unpack"
    e = rel(gw, w.grad)
    assert e < 1e-2, f"conv s2 wgrad rel err {e}"
    gw2, _ = ops.conv_wgrad(g, x_cl, dy_cl, flags=_lib.CONV_NO_SLIDE)
    assert rel(gw, gw2) < 1e-4, "sliding vs streaming wgrad"


@pytest.mark.parametrize("dims,cin,cout,s,n,sp", [(3, 1, 16, 2, 2, (16, 20, 24)), (3, 1, 32, 2, 1, (8, 16, 16)),
                                                  (2, 1, 64, 2, 2, (32, 40)), (2, 3, 16, 1, 1, (17, 19))])
def test_first_layer_im2col(dims, cin, cout, s, n, sp):
    """Small-Cin layer as im2col + 1x1x1 tcgen05 conv: im2col bit-exact against F.unfold-style
    gathering, fprop / wgrad within 1e-2 of torch fp32 on the same bf16-rounded inputs."""
    dtype = torch.bfloat16
    torch.manual_seed(91)
    g = ConvGeom(dims, cin, cout, 3, s, False)
    g1 = ops.col_geom(g)
    assert g1 is not None and g1.cin == cin * 3 ** dims
    w = q(torch.randn(cout, cin, *(3,) * dims) * 0.2, dtype).requires_grad_(True)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype)
    y_ref = ref_conv(g, x, w, b)
    dy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(dy)
    x_cl = cl_dev(x, dtype)
    col = ops.im2col(g, x_cl)
    # reference im2col: column ci*taps + tap of output voxel o = x[o*s - 1 + tap, ci]
    xp = F.pad(x, (1, 1) * dims)
    cols = []
    for ci in range(cin):
        for tap in range(3 ** dims):
            k = [(tap // 3 ** (dims - 1 - i)) % 3 for i in range(dims)]
            sl = [slice(None), ci] + [slice(k[i], k[i] + s * (y_ref.shape[2 + i] - 1) + 1, s) for i in range(dims)]
            cols.append(xp[tuple(sl)])
    col_ref = torch.stack(cols, dim=1)
    assert torch.equal(nc_cpu(col, dims), col_ref), "im2col differs"
    full = col.as_strided(col.shape[:-1] + ((g1.cin + 15) // 16 * 16,), col.stride())
    assert float(full[..., g1.cin:].abs().max()) == 0.0
    wdev = w.detach().to(DEV).reshape(cout, g1.cin, *(1,) * dims)
    y_cl = ops.alloc_activation(n, col.shape[1:4], cout, dtype, DEV)
    ops.conv_fprop(g1, col, ops.pack_weight(g1, _lib.W_CONV_FPROP, wdev, dtype), b.to(DEV), y_cl)
    assert rel(nc_cpu(y_cl, dims), y_ref.detach()) < 1e-2
    dy_cl = ops.alloc_like(y_cl)
    dy_cl.copy_((dy.unsqueeze(2) if dims == 2 else dy).permute(0, 2, 3, 4, 1))
    gw, gb = ops.conv_wgrad(g1, col, dy_cl)
    assert rel(gw.view(w.shape), w.grad) < 1e-2
    assert rel(gb, dy.sum(dim=[0] + list(range(2, 2 + dims)))) < 1e-2


NORM_CASES = [(2, 16, (6, 8, 10)), (1, 10, (8, 8, 12)), (2, 64, (4, 4, 4)), (3, 32, (1, 12, 20)),
              (1, 256, (2, 3, 4)), (2, 7, (3, 5, 7)), (1, 16, (20, 16, 16)), (2, 32, (18, 16, 16)),
              # cluster backward (1024 < voxels <= 32768): 2 / 4 / 8 CTAs per channel group, ragged split
              (2, 64, (16, 16, 16)), (1, 32, (32, 32, 32)), (1, 24, (9, 11, 13)), (2, 128, (8, 8, 8))]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,c,sp", NORM_CASES)
def test_instnorm_prelu(n, c, sp, dtype):
    torch.manual_seed(7)
    x = q(torch.randn(n, c, *sp) * 1.7 + 0.4, dtype).requires_grad_(True)
    alpha = torch.tensor([0.25], requires_grad=True)
    res = q(torch.randn(n, c, *sp), dtype)
    y_ref = F.prelu(F.instance_norm(x, eps=1e-5), alpha) + res
    dy = q(torch.randn_like(y_ref), dtype)
    y_ref.backward(dy)

    x_cl = cl_dev(x.detach(), dtype, pad_c=8, c_off=0)
    mean, rstd = ops.instnorm_stats(x_cl)
    xd = x.detach().double()
    m_ref = xd.mean(dim=(2, 3, 4)).reshape(-1)
    r_ref = (1.0 / torch.sqrt(xd.var(dim=(2, 3, 4), unbiased=False) + 1e-5)).reshape(-1)
    assert rel(mean, m_ref) < 1e-5 and rel(rstd, r_ref) < 1e-5
    a_dev = alpha.detach().to(DEV)
    y_cl = torch.empty((n, *sp, c), dtype=dtype, device=DEV)
    ops.instnorm_prelu_fwd(x_cl, mean, rstd, a_dev, y_cl, cl_dev(res, dtype))
    tol = TOL[dtype]
    assert rel(nc_cpu(y_cl, 3), y_ref.detach()) < tol
    dx_cl = cl_dev(torch.zeros(n, c, *sp), dtype, pad_c=8, c_off=8)
    dalpha = ops.instnorm_prelu_bwd(x_cl, mean, rstd, a_dev, cl_dev(dy, dtype), dx_cl)
    assert rel(nc_cpu(dx_cl, 3), x.grad) < (tol if dtype == torch.float32 else 2e-2)
    assert abs(dalpha.item() - alpha.grad.item()) < (1e-4 if dtype == torch.float32 else 1e-2) * max(1.0, abs(alpha.grad.item()))


PARTIALS_CASES = [
    # cin, cout, stride, transposed, n, input spatial
    (16, 16, 1, False, 2, (12, 40, 24)),    # stride-1 sliding kernel
    (16, 16, 1, False, 2, (10, 21, 64)),    # line-tiled kernel: per-warp partial statistics per item
    (10, 10, 1, False, 1, (7, 6, 128)),
    (32, 32, 1, False, 1, (8, 32, 32)),
    (32, 10, 2, True, 2, (8, 24, 40)),      # ConvTranspose lo->hi, 10 classes padded to 16
    (16, 32, 2, False, 1, (16, 48, 64)),    # stride-2 conv hi->lo
    (64, 64, 1, False, 2, (8, 8, 8)),       # deep layer on the streaming kernel (few partials)
    (128, 256, 1, False, 1, (6, 6, 6)),
]


@pytest.mark.parametrize("with_res", [False, True])
@pytest.mark.parametrize("cin,cout,s,tr,n,sp", PARTIALS_CASES)
def test_conv_partials_instnorm_prelu(cin, cout, s, tr, n, sp, with_res):
    """Convolution -> per-CTA partial statistics -> InstanceNorm+PReLU that finalises them itself,
    against torch fp32 on the same bf16-rounded inputs (1e-2) and against the two-launch path
    (conv_fprop_stats + instnorm_prelu_fwd): same conv output bit for bit, statistics within 1e-6."""
    dtype = torch.bfloat16
    torch.manual_seed(991)
    g = ConvGeom(3, cin, cout, 3, s, tr)
    w = q(torch.randn((cin, cout, 3, 3, 3) if tr else (cout, cin, 3, 3, 3)) * (2.0 / (cin * 27)) ** 0.5, dtype)
    b = torch.randn(cout)
    x = q(torch.randn(n, cin, *sp), dtype)
    alpha = torch.tensor([0.3])
    c_ref = ref_conv(g, x, w, b)
    res = q(torch.randn_like(c_ref), dtype) if with_res else None

    def dev(t_nc):
        out = ops.alloc_activation(t_nc.shape[0], tuple(t_nc.shape[2:]), t_nc.shape[1], dtype, DEV)
        out.copy_(t_nc.permute(0, 2, 3, 4, 1))
        return out

    x_cl = dev(x)
    res_cl = dev(res) if with_res else None
    wp = ops.pack_weight(g, _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP, w.to(DEV), dtype)
    c1 = ops.alloc_activation(n, tuple(c_ref.shape[2:]), cout, dtype, DEV)
    handle = ops.conv_fprop_partials(g, x_cl, wp, b.to(DEV), c1)
    assert handle is not None, "layer did not run on a kernel with fused statistics"
    a1 = ops.alloc_like(c1)
    mean1, rstd1 = ops.instnorm_prelu_fwd_partials(c1, handle, alpha.to(DEV), a1, res_cl)
    c2, a2 = ops.alloc_like(c1), ops.alloc_like(c1)
    mean2, rstd2 = ops.conv_fprop_stats(g, x_cl, wp, b.to(DEV), c2)
    ops.instnorm_prelu_fwd(c2, mean2, rstd2, alpha.to(DEV), a2, res_cl)
    assert torch.equal(c1, c2)
    assert mean1.shape == mean2.shape
    assert rel(mean1, mean2) < 1e-6 and rel(rstd1, rstd2) < 1e-6
    assert rel(a1, a2) < 1e-3
    # the conv output is stored in bf16; the statistics come from the fp32 accumulators
    cd = c_ref.double()
    m_ref = cd.mean(dim=(2, 3, 4))
    r_ref = 1.0 / torch.sqrt(cd.var(dim=(2, 3, 4), unbiased=False) + 1e-5)
    assert rel(mean1.view(n, -1)[:, :cout], m_ref) < 2e-3 and rel(rstd1.view(n, -1)[:, :cout], r_ref) < 2e-3
    y_ref = F.prelu(F.instance_norm(c_ref, eps=1e-5), alpha)
    if with_res:
        y_ref = y_ref + res
    assert rel(nc_cpu(a1, 3), y_ref) < 1e-2
    if cout % 16:  # channel padding stays zero
        full = a1.as_strided(a1.shape[:-1] + ((cout + 15) // 16 * 16,), a1.stride())
        assert float(full[..., cout:].abs().max()) == 0.0


# ---- Focal / CrossEntropy through the shared softmax pass ----------------------------------------
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("c,sp", [(10, (6, 10, 12)), (4, (1, 9, 11)), (10, (16, 32, 32))])
def test_focal_and_cross_entropy_vs_oracle(c, sp, dtype):
    """Focal (monai 0.3, one-hot target), CrossEntropy and class-weighted CrossEntropy values and logit
    gradients from ONE fused softmax pass, against the oracle / torch on the same rounded logits."""
    torch.manual_seed(5)
    n = 2
    z = q(torch.randn(n, c, *sp) * 2.0, dtype).requires_grad_(True)
    lab = torch.randint(0, c, (n, *sp))
    lab[0, 0, 0, :3] = 0
    onehot = O.one_hot(lab.unsqueeze(1), c)
    w = torch.rand(c) + 0.05
    tol = 1e-4 if dtype == torch.float32 else 2e-3
    zd = cl_dev(z.detach(), dtype)                      # channels-last device logits
    zin = zd.permute(0, 4, 1, 2, 3).requires_grad_(True)
    cases = {
        "focal_mean": (lambda: O.FocalLoss(reduction="mean")(z, onehot),
                       lambda x: losses.FocalLoss(reduction="mean")(x, lab.to(DEV).unsqueeze(1))),
        "focal_none": (lambda: O.FocalLoss(reduction="none")(z, onehot).square().sum(),
                       lambda x: losses.FocalLoss(reduction="none")(x, lab.to(DEV).unsqueeze(1)).square().sum()),
        "ce": (lambda: F.cross_entropy(z, lab), lambda x: losses.CrossEntropyLoss()(x, lab.to(DEV))),
        "wce": (lambda: F.cross_entropy(z, lab, weight=w),
                lambda x: losses.CrossEntropyLoss(weight=w)(x, lab.to(DEV))),
    }
    for name, (ref_fn, dev_fn) in cases.items():
        z.grad = None
        lr = ref_fn()
        lr.backward()
        zin.grad = None
        ld = dev_fn(zin)
        ld.backward()
        assert abs(ld.item() - lr.item()) < tol * max(1.0, abs(lr.item())), f"{name}: {ld.item()} vs {lr.item()}"
        e = rel(zin.grad.float().cpu(), z.grad)
        assert e < (1e-4 if dtype == torch.float32 else 1e-2), f"{name} gradient rel err {e}"


def test_multiple_loss_wrapper_shared_pass():
    """MultipleLossWrapper with Dice + Focal + CrossEntropy (one fused pass) == the oracle's separate losses,
    with and without the AnatomyNet missing-annotation weighting."""
    torch.manual_seed(11)
    n, c, sp = 3, 10, (4, 12, 12)
    z = torch.randn(n, c, *sp) * 1.5
    lab = torch.randint(0, c, (n, *sp))
    ind = torch.ones(n, c - 1)
    ind[0, 2] = 0
    ind[1, 5] = 0
    for excl in (False, True):
        names = ["CrossEntropy", "Dice", "Focal"]
        got = losses.MultipleLossWrapper(names, exclude_missing=excl)(z.to(DEV), lab.to(DEV), ind.to(DEV))
        ref = O.MultipleLossWrapper(names, exclude_missing=excl)(z, lab, ind)
        for k in names:
            assert abs(got[k].item() - ref[k].item()) < 1e-4 * max(1.0, abs(ref[k].item())), (excl, k, got[k].item(), ref[k].item())


# ---- softmax + Dice ---------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["dense", "sparse"])
@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int64])
def test_dice_loss_golden_fp32(golden, tag, label_dtype):
    logits = torch.from_numpy(golden["dice_logits"]).to(DEV).requires_grad_(True)
    lab = torch.from_numpy(golden[f"dice_lab_{tag}"]).to(DEV, label_dtype).unsqueeze(1)
    fx = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="mean")
    v = fx(logits, lab)
    v.backward()
    assert abs(v.item() - float(golden[f"dice_{tag}_mean"])) < 1e-5      # north_star: Dice loss within 1e-3
    assert rel(logits.grad, torch.from_numpy(golden[f"dice_{tag}_grad"])) < 1e-4
    fxn = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="none")
    np.testing.assert_allclose(fxn(logits.detach(), lab).cpu().numpy(), golden[f"dice_{tag}_none"],
                               rtol=1e-4, atol=1e-6)


def test_dice_loss_2d_and_missing_mask(golden):
    fxn = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, reduction="none")
    v = fxn(torch.from_numpy(golden["dice2d_logits"]).to(DEV),
            torch.from_numpy(golden["dice2d_lab"]).to(DEV).unsqueeze(1))
    np.testing.assert_allclose(v.cpu().numpy(), golden["dice2d_none"], rtol=1e-4, atol=1e-6)
    wrap = losses.MultipleLossWrapper(["Dice"], exclude_missing=True)
    logits = torch.from_numpy(golden["dice_logits"]).to(DEV)
    lab = torch.from_numpy(golden["dice_lab_sparse"]).to(DEV)
    for ind_key, out_key in (("indicator", "missing_dice"), ("indicator_inf", "missing_dice_inf")):
        out = wrap(logits, lab, torch.from_numpy(golden[ind_key]).to(DEV))
        assert abs(out["Dice"].item() - float(golden[out_key])) < 1e-5


@pytest.mark.parametrize("wt", ["square", "simple", "uniform"])
@pytest.mark.parametrize("tag", ["dense", "sparse"])
def test_generalized_dice_golden(golden, wt, tag):
    """losses.GeneralizedDiceLoss vs the reference's own capstone/models/temp.py class: default "square" weights,
    "simple", and the inf -> max rule (sparse labels: classes absent from a sample), value / (B, 9) matrix / gradient."""
    logits = torch.from_numpy(golden["dice_logits"]).to(DEV).requires_grad_(True)
    lab = torch.from_numpy(golden[f"dice_lab_{tag}"]).to(DEV).unsqueeze(1)
    key = f"gdl_{wt}_{tag}" if wt != "uniform" else f"dice_{tag}"
    v = losses.GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True, w_type=wt)(logits, lab)
    v.backward()
    assert abs(v.item() - float(golden[f"{key}_mean"])) < 1e-5
    assert rel(logits.grad, torch.from_numpy(golden[f"{key}_grad"])) < 1e-4
    fxn = losses.GeneralizedDiceLoss(include_background=False, to_onehot_y=True, softmax=True, w_type=wt,
                                     reduction="none")
    np.testing.assert_allclose(fxn(logits.detach(), lab).cpu().numpy(), golden[f"{key}_none"], rtol=1e-4, atol=2e-6)
    if wt == "square":  # the wrapper the LOSSES table builds (reference default w_type)
        w = losses.MultipleLossWrapper(["GeneralizedDice"])(logits.detach(), lab[:, 0])
        assert abs(w["GeneralizedDice"].item() - float(golden[f"{key}_mean"])) < 1e-5


def test_boundary_loss_golden(golden):
    """Boundary loss (sixth sum of the shared softmax pass) vs the reference's own BoundaryLossWrapper /
    MultipleLossWrapper(["Boundary"], exclude_missing=True): values, (B, 9) matrix, logit gradients."""
    logits = torch.from_numpy(golden["boundary_logits"]).to(DEV).requires_grad_(True)
    dist = torch.from_numpy(golden["boundary_dist"]).to(DEV)
    v = losses.BoundaryLossWrapper("mean")(logits, dist)
    v.backward()
    assert abs(v.item() - float(golden["boundary_mean"])) < 1e-7
    assert rel(logits.grad, torch.from_numpy(golden["boundary_grad"])) < 1e-4
    none = losses.BoundaryLossWrapper("none")(logits.detach(), dist)
    np.testing.assert_allclose(none.cpu().numpy(), golden["boundary_none"], rtol=1e-4, atol=1e-9)
    lab = torch.from_numpy(golden["boundary_lab"]).to(DEV)
    ind = torch.from_numpy(golden["indicator"]).to(DEV)
    logits.grad = None
    out = losses.MultipleLossWrapper(["Boundary"], exclude_missing=True)(logits, lab, ind, dist)
    out["Boundary"].backward()
    assert abs(out["Boundary"].item() - float(golden["boundary_missing"])) < 1e-7
    assert rel(logits.grad, torch.from_numpy(golden["boundary_missing_grad"])) < 1e-4
    with pytest.raises(AssertionError):
        losses.MultipleLossWrapper(["Boundary"])(logits, lab, ind)  # reference: distance maps are required


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_boundary_shares_the_softmax_pass(dtype):
    """Boundary + Dice + Focal + GeneralizedDice + CrossEntropy through ONE fused pass (3-D, with the 16-channel
    padded bf16 layout) == the oracle's separate losses, values and the gradient of their sum."""
    torch.manual_seed(21)
    n, c, sp = 2, 10, (8, 16, 12)
    z = q(torch.randn(n, c, *sp) * 1.5, dtype).requires_grad_(True)
    lab = torch.randint(0, c, (n, *sp))
    dist = torch.randn(n, c - 1, *sp) * 0.05
    names = ["Boundary", "CrossEntropy", "Dice", "Focal", "GeneralizedDice"]
    ref = O.MultipleLossWrapper(names)(z, lab, None, dist)
    torch.stack(list(ref.values())).sum().backward()
    zd = cl_dev(z.detach(), dtype, pad_c=6 if dtype == torch.bfloat16 else 0, c_off=0)
    zin = zd.permute(0, 4, 1, 2, 3).requires_grad_(True)
    got = losses.MultipleLossWrapper(names)(zin, lab.to(DEV), None, dist.to(DEV))
    torch.stack(list(got.values())).sum().backward()
    tol = 1e-4 if dtype == torch.float32 else 2e-3
    for k in names:
        assert abs(got[k].item() - ref[k].item()) < tol * max(1.0, abs(ref[k].item())), (k, got[k].item(), ref[k].item())
    assert rel(zin.grad.float().cpu(), z.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,sp,pad", [(2, (16, 24, 20), 0), (1, (40, 48, 32), 6), (3, (1, 64, 96), 0)])
def test_dice_loss_vs_oracle(n, sp, pad, dtype):
    torch.manual_seed(3)
    c = 10
    logits = q(torch.randn(n, c, *sp) * 3, dtype).requires_grad_(True)
    lab = torch.randint(0, c, (n, *sp))
    lab[0][lab[0] == 4] = 0  # an absent class
    ref = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(logits, lab.unsqueeze(1))
    ref.backward()
    cl = cl_dev(logits.detach(), dtype, pad_c=pad, c_off=0)
    sums = ops.softmax_dice_sums(cl, lab.to(DEV, torch.uint8))
    s = sums[:, 1:]
    f = 1.0 - (2.0 * s[..., 0] + 1e-5) / (s[..., 1] + s[..., 2] + 1e-5)
    assert abs(f.mean().item() - ref.item()) < 1e-3
    # G is an exact integer count
    gt = torch.stack([torch.bincount(lab[i].reshape(-1), minlength=c) for i in range(n)]).float()
    assert torch.equal(sums[..., 1].cpu(), gt)
    # backward through the public module
    inp = torch.from_numpy(np.ascontiguousarray(logits.detach().numpy())).to(DEV, dtype).requires_grad_(True)
    v = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(inp, lab.to(DEV).unsqueeze(1))
    v.backward()
    assert abs(v.item() - ref.item()) < 1e-3
    assert rel(inp.grad.float(), logits.grad) < (1e-4 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("label_dtype", [torch.uint8, torch.int64])
@pytest.mark.parametrize("n,sp", [(2, (16, 24, 20)), (1, (5, 7, 6)), (3, (1, 33, 47)), (2, (40, 64, 64))])
def test_dice_ring_kernels_and_fused_metric(n, sp, label_dtype):
    """Production layout (bf16, 16-channel rows, 10 classes): the bulk-copy staged forward / backward kernels equal
    the direct-load kernels (B200SEG_DICE_NO_RING=1) and the oracle, for ragged voxel counts too; the metric counts
    that ride on the forward pass equal the stand-alone argmax + count kernel exactly."""
    torch.manual_seed(17)
    c = 10
    logits = q(torch.randn(n, c, *sp) * 3, torch.bfloat16).requires_grad_(True)
    lab = torch.randint(0, c, (n, *sp))
    lab[0][lab[0] == 7] = 0
    ref = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(logits, lab.unsqueeze(1))
    ref.backward()
    cl = cl_dev(logits.detach(), torch.bfloat16, pad_c=6, c_off=0)       # ld = 16
    assert ops.cl_info(cl)[5] == 16
    labd = lab.to(DEV, label_dtype)
    lib = _lib.load()
    sums = ops.softmax_dice_sums(cl, labd)
    assert lib.b200seg_last_launch() in (b"dice_sums_final",)
    sums_m, counts = ops.softmax_dice_metric_sums(cl, labd)
    assert lib.b200seg_last_launch() == b"dice_metric_final"            # the one-pass kernel ran
    gi, gp = torch.rand(n, c, device=DEV) - 0.5, torch.rand(n, c, device=DEV) - 0.5
    dz = ops.softmax_dice_bwd(cl, labd, gi, gp)
    os.environ["B200SEG_DICE_NO_RING"] = "1"
    try:
        sums_d = ops.softmax_dice_sums(cl, labd)
        dz_d = ops.softmax_dice_bwd(cl, labd, gi, gp)
    finally:
        del os.environ["B200SEG_DICE_NO_RING"]
    torch.testing.assert_close(sums, sums_d, rtol=1e-5, atol=1e-4)
    assert torch.equal(sums_m, sums)
    assert torch.equal(dz, dz_d)                                         # same arithmetic, same order
    full = dz.as_strided(dz.shape[:-1] + (16,), dz.stride())
    assert float(full[..., c:].abs().max()) == 0.0                       # channel padding stays zero
    _, counts_ref = ops.argmax_dice_counts(cl, labd, want_pred=False)
    assert torch.equal(counts, counts_ref)
    s = sums[:, 1:]
    f = 1.0 - (2.0 * s[..., 0] + 1e-5) / (s[..., 1] + s[..., 2] + 1e-5)
    assert abs(f.mean().item() - ref.item()) < 1e-3
    # public module: loss + metric from one pass, gradient through the ring backward
    fx = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, with_metric=True)
    inp = cl.permute(0, 4, 1, 2, 3).requires_grad_(True)
    v = fx(inp, labd.unsqueeze(1))
    v.backward()
    assert abs(v.item() - ref.item()) < 1e-3
    assert rel(inp.grad.float(), logits.grad) < 1e-2
    dm, dpc = metrics.dice_from_counts(fx.metric_counts)
    dm_ref, dpc_ref = O.dice_metric(O.squash_predictions(logits.detach()), lab)
    mism = (metrics.squash_predictions(inp.detach()).cpu() != O.squash_predictions(logits.detach())).sum().item()
    if mism == 0:
        np.testing.assert_allclose(dpc.cpu().numpy(), dpc_ref.numpy(), rtol=1e-6, atol=1e-7)


def test_dice_ring_kernels_full_size_race_free():
    """cfg5's loss tensors (4 x 160^3, 28 ring refills per CTA, up to 7 CTAs per SM): the staged kernels equal the
    direct-load kernels bit for bit, run after run.  Regression test for the write-after-read hazard r2 found here (a
    bulk-copy refill issued before the previous rows had been CONSUMED overtook the shared-memory loads: dlogits
    differed from run to run by up to 0.25 at this size, while 2 x 128^3 was clean)."""
    torch.manual_seed(23)
    n, p_ = 4, 160
    z = ops.alloc_activation(n, (p_, p_, p_), 10, torch.bfloat16, DEV)
    z.copy_(torch.randn(z.shape, device=DEV) * 3)
    lab = torch.randint(0, 10, (n, p_, p_, p_), device=DEV, dtype=torch.uint8)
    gi, gp = torch.rand(n, 10, device=DEV) - 0.5, torch.rand(n, 10, device=DEV) - 0.5
    os.environ["B200SEG_DICE_NO_RING"] = "1"
    try:
        dz_ref = ops.softmax_dice_bwd(z, lab, gi, gp)
        sums_ref = ops.softmax_dice_sums(z, lab)
        _, counts_ref = ops.argmax_dice_counts(z, lab, want_pred=False)
    finally:
        del os.environ["B200SEG_DICE_NO_RING"]
    dz = ops.alloc_like(z)
    for _ in range(4):
        dz.fill_(7.0)
        ops.softmax_dice_bwd(z, lab, gi, gp, dlogits=dz)
        assert _lib.load().b200seg_last_launch() == b"softmax_dice_bwd_ring"
        assert torch.equal(dz, dz_ref)
        sums, counts = ops.softmax_dice_metric_sums(z, lab)
        torch.testing.assert_close(sums, sums_ref, rtol=1e-5, atol=1e-2)
        assert torch.equal(counts, counts_ref)
    first = ops.softmax_dice_metric_sums(z, lab)[0]
    assert all(torch.equal(first, ops.softmax_dice_metric_sums(z, lab)[0]) for _ in range(3))


def test_dice_metric_fused_other_layouts(golden):
    """fp32 check mode / unpadded rows: b200seg_softmax_dice_metric_fwd falls back to loss kernel + metric kernel and
    gives the reference's metric (golden, bit-exact argmax) and loss."""
    logits = torch.from_numpy(golden["dice_logits"]).to(DEV)
    lab = torch.from_numpy(golden["dice_lab_sparse"]).to(DEV)
    fx = losses.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, with_metric=True)
    v = fx(logits, lab.unsqueeze(1))
    assert abs(v.item() - float(golden["dice_sparse_mean"])) < 1e-5
    dm, dpc = metrics.dice_from_counts(fx.metric_counts)
    np.testing.assert_allclose(dpc.cpu().numpy(), golden["metric_sparse_per_class"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dm.item(), golden["metric_sparse_mean"], rtol=1e-6, atol=1e-7)


def test_label_maps_golden(golden):
    pred = metrics.squash_predictions(torch.from_numpy(golden["dice_logits"]).to(DEV))
    assert np.array_equal(pred.cpu().numpy().astype(np.uint8), golden["argmax"])          # bit-exact
    tie = metrics.squash_predictions(torch.from_numpy(golden["tie_logits"]).to(DEV))
    assert np.array_equal(tie.cpu().numpy().astype(np.uint8), golden["tie_argmax"])
    m3 = metrics.squash_masks(torch.from_numpy(golden["masks"]).to(DEV))
    assert np.array_equal(m3.cpu().numpy(), golden["squash3d"])
    m2 = metrics.squash_masks(torch.from_numpy(golden["masks2d"]).to(DEV))
    assert np.array_equal(m2.cpu().numpy(), golden["squash2d"])


def test_argmax_large_bitexact_vs_torch():
    torch.manual_seed(11)
    logits = torch.randn(2, 10, 32, 48, 40)
    ref = torch.softmax(logits, dim=1).argmax(dim=1)
    got = metrics.squash_predictions(logits.to(DEV))
    assert torch.equal(got.cpu(), ref)


def test_dice_metric_golden(golden):
    wrap = metrics.DiceMetricWrapper()
    pred = torch.from_numpy(golden["argmax"]).to(DEV)
    for tag in ("dense", "sparse"):
        for ldt in (torch.uint8, torch.int64):
            dm, dpc = wrap(pred, torch.from_numpy(golden[f"dice_lab_{tag}"]).to(DEV, ldt))
            np.testing.assert_allclose(dpc.cpu().numpy(), golden[f"metric_{tag}_per_class"], rtol=1e-6, atol=1e-7)
            np.testing.assert_allclose(dm.item(), golden[f"metric_{tag}_mean"], rtol=1e-6, atol=1e-7)
    dm, dpc = wrap(torch.from_numpy(golden["metric_noisy_pred"]).to(DEV),
                   torch.from_numpy(golden["metric_noisy_target"]).to(DEV))
    np.testing.assert_allclose(dpc.cpu().numpy(), golden["metric_noisy_per_class"], rtol=1e-6, atol=1e-7)
    # fused logits -> metric path equals the two-step path
    logits = torch.from_numpy(golden["dice_logits"]).to(DEV)
    dm2, dpc2 = wrap.from_logits(logits, torch.from_numpy(golden["dice_lab_sparse"]).to(DEV))
    np.testing.assert_allclose(dpc2.cpu().numpy(), golden["metric_sparse_per_class"], rtol=1e-6, atol=1e-7)


def test_dice_counts_checksum_large():
    """Size-independent property at full size: counts sum to the voxel count; tp <= min(pred, target)."""
    torch.manual_seed(5)
    n, sp = 2, (96, 96, 96)
    pred = torch.randint(0, 10, (n, *sp), dtype=torch.uint8, device=DEV)
    tgt = torch.randint(0, 10, (n, *sp), dtype=torch.uint8, device=DEV)
    counts = ops.label_dice_counts(pred, tgt, 10)
    vox = sp[0] * sp[1] * sp[2]
    assert torch.equal(counts[..., 1].sum(1).cpu(), torch.full((n,), vox))
    assert torch.equal(counts[..., 2].sum(1).cpu(), torch.full((n,), vox))
    assert torch.equal(counts[..., 0].sum(1).cpu(), (pred == tgt).reshape(n, -1).sum(1).cpu())
    assert bool((counts[..., 0] <= torch.minimum(counts[..., 1], counts[..., 2])).all())


def test_hu_windowing_golden(golden):
    hu = torch.from_numpy(golden["hu"]).to(DEV)  # (32, 40, 1) int16
    out = transforms.window_normalize(hu[..., 0], ("brain", "soft_tissue", "bone"))
    ref = O.window_normalize(golden["hu"][..., 0], ("brain", "soft_tissue", "bone"))
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=1e-6)
    # un-normalised window equals the reference's apply_window (to float32 rounding)
    raw = transforms.window_normalize(hu[..., 0], ("soft_tissue",), mean=[0.0], std=[1.0])
    np.testing.assert_allclose(raw.cpu().numpy()[..., 0], golden["window_soft_tissue"][..., 0].astype(np.float32),
                               rtol=0, atol=1e-7)


def test_error_paths():
    lib = _lib.load()
    g = ConvGeom(3, 16, 16, 3, 1, False)
    x = torch.zeros(1, 4, 4, 4, 16, device=DEV)
    y = torch.zeros(1, 4, 4, 5, 16, device=DEV)
    with pytest.raises(ValueError):
        ops.conv_fprop(g, x, ops.pack_weight(g, 0, torch.zeros(16, 16, 3, 3, 3, device=DEV), torch.float32), None, y)
    with pytest.raises(RuntimeError):
        ops.cl_info(torch.zeros(1, 2, 2, 2, 4))  # CPU tensor: no fallback
    with pytest.raises(TypeError):
        ops.dtype_code(torch.float16)
    assert lib.b200seg_check_device(0) == 0
