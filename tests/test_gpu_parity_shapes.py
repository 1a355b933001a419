"""Parity AT THE BENCHMARKED SHAPES -- `pytest -m gpu`.

Kernel dispatch is size dependent (sliding-window kernels need >= 8 x 512 voxels, deferred statistics switch at
64^3, TMEM ring wrap, grid segmentation), so the shapes BASELINE.json names are run whole through
``B.UNet`` + ``DiceLoss`` and compared with the CPU oracle on the same seeded inputs and weights:

* fp32 check mode: logits 1e-4, Dice 1e-5, label map bit-exact (away from exact ties), every layer's forward and
  backward recomputed from ITS OWN inputs with torch fp32 within 1e-4;
* bf16: Dice 1e-3, every layer's forward output and backward (input gradient of InstanceNorm+PReLU, weight
  gradient, PReLU slope) recomputed from ITS OWN bf16 inputs with torch fp32 within **1e-2** (the north-star
  bound, layer by layer on identical inputs), and the END-TO-END drift (logits and parameter gradients vs the fp32
  oracle) calibrated against stock torch bf16 (cuDNN autocast, channels_last) on the same GPU: ours must not drift
  more than torch's own bf16 path does (x CAL_FACTOR).  The measured table is written to gpurun_out/ for profiles/.
"""
import json
import os

import pytest
import torch
import torch.nn.functional as F

import ct_image_segmentation_b200 as B
from ct_image_segmentation_b200 import ops
from ct_image_segmentation_b200.unet import Convolution
from oracle import monai_ref as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
CAL_FACTOR = 1.5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SMALL, WIDE, NET2D = [16, 32, 64, 128, 256], [32, 64, 128, 256, 512], [64, 128, 256, 512, 1024]
SHAPES = {
    # BASELINE.json configs[2]: 128^3 patches (1 and 2 per GPU), configs[1]: 96^3 x 2, configs[4] net at 64^3,
    # configs[0]: the 2-D 64-1024 net (at 256^2: the oracle's 512^2 x 4 step takes minutes on the host)
    "cfg3_128x1": (3, SMALL, (1, 1, 128, 128, 128)),
    "cfg3_128x2": (3, SMALL, (2, 1, 128, 128, 128)),
    "cfg2_96x2": (3, SMALL, (2, 1, 96, 96, 96)),
    "cfg5net_64x1": (3, WIDE, (1, 1, 64, 64, 64)),
    "cfg1net_256x2": (2, NET2D, (2, 1, 256, 256)),
}


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make_pair(dims, channels, dtype, seed=12342):
    torch.manual_seed(seed)
    ref = O.UNet(dims, 1, 10, channels, [2, 2, 2, 2], num_res_units=2)
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith("act.weight"):
                p.uniform_(0.1, 0.4)
    net = B.UNet(dims, 1, 10, channels, [2, 2, 2, 2], num_res_units=2, dtype=dtype)
    net.load_state_dict(ref.state_dict())
    return ref, net.to(DEV)


def blob_labels(n, sp, seed=0):
    g = torch.Generator().manual_seed(seed)
    lab = torch.zeros(n, *sp, dtype=torch.int64)
    for i in range(n):
        for c in range(1, 10):
            lo = [int(torch.randint(0, max(1, s - s // 3), (1,), generator=g)) for s in sp]
            sl = tuple(slice(l, l + max(2, s // 4)) for l, s in zip(lo, sp))
            lab[(i, *sl)] = c
    return lab


def _nc(t, dims):
    t = t.float().cpu().permute(0, 4, 1, 2, 3)
    return t.squeeze(2) if dims == 2 else t


def _conv(g, x, w, b=None):
    p = (g.kernel - 1) // 2
    if g.transposed:
        f = F.conv_transpose2d if g.dims == 2 else F.conv_transpose3d
        return f(x, w, b, stride=g.stride, padding=p, output_padding=g.stride - 1)
    f = F.conv2d if g.dims == 2 else F.conv3d
    return f(x, w, b, stride=g.stride, padding=p)


def ref_im2col(x, g):
    """(N, cin*taps, *out) with column ci*taps + tap = x[ci] at the tap's input position (zero outside)."""
    k, s, dims = g.kernel, g.stride, g.dims
    if dims == 2:
        x = x.unsqueeze(2)
    p = (k - 1) // 2
    xp = F.pad(x, (p, p, p, p) + ((p, p) if dims == 3 else (0, 0)))
    n, c, D, H, W = x.shape
    od = (D + 2 * p - k) // s + 1 if dims == 3 else 1
    oh, ow = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    cols = []
    for ci in range(c):
        for kd in range(k if dims == 3 else 1):
            for kh in range(k):
                for kw in range(k):
                    cols.append(xp[:, ci, kd:kd + (od - 1) * s + 1:s, kh:kh + (oh - 1) * s + 1:s,
                                   kw:kw + (ow - 1) * s + 1:s])
    out = torch.stack(cols, 1)
    return out.squeeze(2) if dims == 2 else out


def wgrad_err(g, xin, w, gy, got, shape, tol=1e-4):
    """Relative error of a weight gradient against torch's on the same (x, dy).  A weight gradient is a sum over up
    to 4 M voxels: torch's own fp32 CPU kernel is then off by a few 1e-4 (r2 measurement: the fp32 check kernels and
    the tcgen05 kernels showed the SAME 3.2e-4 / 7.4e-4 / 1.4e-4 against it at 128^3 x 1 / x 2 / 96^3 x 2) -- where
    the fp32 reference disagrees, a float64 reference arbitrates."""
    wz = torch.zeros_like(w).requires_grad_(True)
    (gw_ref,) = torch.autograd.grad(_conv(g, xin, wz), wz, gy)
    e = rel(got, gw_ref.reshape(shape))
    if e >= tol:
        wz = torch.zeros_like(w, dtype=torch.float64).requires_grad_(True)
        (gw_ref,) = torch.autograd.grad(_conv(g, xin.double(), wz), wz, gy.double())
        e = rel(got, gw_ref.reshape(shape))
    return e


def layer_local_check(net, dims, x, lab, dtype, tol):
    """Every convolution's forward (conv, InstanceNorm + PReLU [+ residual]) and backward recomputed by torch fp32
    from the layer's OWN inputs as the GPU run saw them -- on the code path the benchmark runs (im2col first layers
    included).  Returns (logits, loss, list of violations, worst errors)."""
    from ct_image_segmentation_b200.unet import ResidualUnit
    saved = {}
    x_cl = ops.to_channels_last(x.to(DEV), dtype)
    out = net._run_forward(x_cl, saved, keep_all=True, use_cols=True)
    lg = ops.from_channels_last(out, dims).detach().requires_grad_(True)
    loss = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(lg, lab.to(DEV).unsqueeze(1))
    loss.backward()
    fwd = {m: dict(s) for m, s in saved.items() if isinstance(m, (Convolution, ResidualUnit))}
    taps = {}
    grads, _ = net._run_backward(saved, ops.to_channels_last(lg.grad, dtype), False, taps)
    names = {m: n for n, m in net.named_modules()}
    bad, worst = [], {"conv": 0.0, "act": 0.0, "in_bwd": 0.0, "wgrad": 0.0, "im2col": 0.0}
    wq = (lambda w: w.to(dtype).float())  # the kernels read the weights in the compute dtype

    def note(kind, name, what, e):
        worst[kind] = max(worst[kind], e)
        if not (e < tol):
            bad.append((name, what, e))

    def conv_in(g, g1, x_or_col, w):
        """(geometry, NC input, weight) the kernel really saw: the layer itself, or its 1x1 form on the im2col."""
        if g1 is None:
            return g, _nc(x_or_col, dims), w
        return g1, _nc(x_or_col, dims), w.reshape(w.shape[0], -1, *([1] * dims))

    n_col = 0
    for m, t in taps.items():
        name = names[m]
        if isinstance(m, ResidualUnit):
            g, w, b = m.res_geom, m.residual.weight.detach().cpu(), m.residual.bias.detach().cpu()
            g1 = t["col_geom"]
            ge, xin, we = conv_in(g, g1, t["col"] if g1 is not None else t["x"], w)
            if t["r"] is not None:
                note("conv", name, "residual conv out", rel(_nc(t["r"], dims), _conv(ge, xin, wq(we), b)))
            gy = _nc(t["g_out"], dims)
            note("wgrad", name, "residual wgrad", wgrad_err(ge, xin, we, gy, grads[m.residual.weight], w.shape))
            gb_ref = gy.sum(dim=[0] + list(range(2, gy.dim())))
            note("wgrad", name, "residual bias grad", rel(grads[m.residual.bias], gb_ref))
            continue
        g, s = m.geom, fwd[m]
        w = m.conv.weight.detach().cpu()
        g1 = s.get("col_geom")
        if g1 is not None:  # the im2col buffer itself, against the layer's real input
            n_col += 1
            src = x_cl if name.startswith("model.0.") else None
            if src is not None:
                note("im2col", name, "im2col", rel(_nc(t["x"], dims), ref_im2col(_nc(src, dims), g)))
        ge, xin, we = conv_in(g, g1, t["x"], w)
        c_ref = _conv(ge, xin, wq(we), m.conv.bias.detach().cpu())
        res = None if s.get("res") is None else _nc(s["res"], dims)
        if s.get("c") is not None:
            note("conv", name, "conv out", rel(_nc(s["c"], dims), c_ref))
            c_gpu = _nc(s["c"], dims)
            a_ref = F.prelu(F.instance_norm(c_gpu, eps=m.norm.eps), m.act.weight.detach().cpu())
            if res is not None:
                a_ref = a_ref + res
            note("act", name, "in+prelu(+res) out", rel(_nc(s["out"], dims), a_ref))
        else:  # conv-only head: bias (+ residual) fused
            note("conv", name, "conv(+res) out", rel(_nc(s["out"], dims), c_ref if res is None else c_ref + res))
        g_out, g_c = t["g_out"].float().cpu(), t["g_c"].float().cpu()
        if t["c"] is not None:
            n, c = g_c.shape[0], g_c.shape[-1]
            mean = t["mean"].cpu().view(n, -1)[:, :c].reshape(n, 1, 1, 1, c)
            rstd = t["rstd"].cpu().view(n, -1)[:, :c].reshape(n, 1, 1, 1, c)
            h = (t["c"].float().cpu() - mean) * rstd
            alpha = m.act.weight.detach().cpu()
            gt = torch.where(h > 0, g_out, alpha * g_out)
            ref_gc = rstd * (gt - gt.mean(dim=(1, 2, 3), keepdim=True) - h * (gt * h).mean(dim=(1, 2, 3), keepdim=True))
            note("in_bwd", name, "in+prelu bwd", rel(g_c, ref_gc))
            terms = torch.where(h > 0, torch.zeros_like(h), g_out * h)
            da = grads[m.act.weight].item()
            if abs(da - terms.double().sum().item()) > 1e-3 * terms.abs().double().sum().item() + 1e-12:
                bad.append((name, "dalpha", da, terms.double().sum().item()))
        gy = _nc(g_c, dims)
        note("wgrad", name, "wgrad", wgrad_err(ge, xin, we, gy, grads[m.conv.weight], w.shape))
    if dtype == torch.bfloat16:
        assert n_col >= 1, "the benchmark path runs the first layer through the im2col kernels"
    return lg.detach(), loss.detach(), bad, worst


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=["fp32check", "bf16"])
@pytest.mark.parametrize("case", list(SHAPES))
def test_benchmark_shape_parity(case, dtype):
    dims, channels, shape = SHAPES[case]
    fp32 = dtype == torch.float32
    ref, net = make_pair(dims, channels, dtype)
    torch.manual_seed(1)
    x = torch.randn(*shape)
    lab = blob_labels(shape[0], shape[2:])
    y_ref = ref(x)
    loss_ref = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y_ref, lab.unsqueeze(1))
    y, loss, bad, worst = layer_local_check(net, dims, x, lab, dtype, 1e-4 if fp32 else 1e-2)
    print(f"\n[{case} {'fp32' if fp32 else 'bf16'}] layer-local worst errors {worst}; logits rel "
          f"{rel(y, y_ref.detach()):.3e}; dice {loss.item():.6f} vs {loss_ref.item():.6f}")
    assert not bad, f"layer-local parity out of tolerance: {bad}"
    assert abs(loss.item() - loss_ref.item()) < (1e-5 if fp32 else 1e-3)     # north_star: Dice within 1e-3
    if fp32:
        assert rel(y, y_ref.detach()) < 1e-4
        lm_ref = O.squash_predictions(y_ref.detach())
        lm = B.squash_predictions(y)
        mism = (lm.cpu() != lm_ref)
        if mism.any():  # bit-exact away from ties that 1e-6 of logit noise can flip
            top2 = torch.softmax(y_ref.detach(), 1).topk(2, dim=1).values
            gap = (top2[:, 0] - top2[:, 1])[mism]
            assert float(gap.max()) < 1e-5, f"{int(mism.sum())} label mismatches, margin up to {float(gap.max())}"
            assert int(mism.sum()) <= 1e-5 * mism.numel() + 2


def _hooked_forward(ref, x):
    taps, hooks = {}, []
    for name, m in ref.named_modules():
        if isinstance(m, O.Convolution):
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: taps.__setitem__(name, o.detach().float())))
    y = ref(x)
    for h in hooks:
        h.remove()
    return y, taps


@pytest.mark.parametrize("case", ["cfg3_128x1", "cfg2_96x2", "cfg5net_64x1"])
def test_bf16_drift_calibrated_against_torch_bf16(case):
    """End-to-end bf16 drift of the whole network, calibrated: the fp32 oracle, stock torch bf16 (autocast +
    channels_last, cuDNN) and this path run on the same GPU, same inputs and weights.  Per Convolution output,
    for the logits and per parameter gradient: err(ours vs fp32) <= CAL_FACTOR * err(torch-bf16 vs fp32) (+ floor)."""
    dims, channels, shape = SHAPES[case]
    ref, net = make_pair(dims, channels, torch.bfloat16)
    torch.manual_seed(1)
    x = torch.randn(*shape)
    lab = blob_labels(shape[0], shape[2:])
    fx = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    old_tf32 = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = ref.to(DEV)
        xg, lg = x.to(DEV), lab.to(DEV)
        # (1) fp32 oracle on the GPU (no TF32)
        y32, taps32 = _hooked_forward(ref, xg)
        fx(y32, lg.unsqueeze(1)).backward()
        g32 = {n: p.grad.detach().clone() for n, p in ref.named_parameters()}
        for p in ref.parameters():
            p.grad = None
        # (2) stock torch bf16
        mf = torch.channels_last_3d if dims == 3 else torch.channels_last
        ref_cl = ref.to(memory_format=mf)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y16, taps16 = _hooked_forward(ref_cl, xg.contiguous(memory_format=mf))
        fx(y16.float(), lg.unsqueeze(1)).backward()
        g16 = {n: p.grad.detach().clone() for n, p in ref_cl.named_parameters()}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old_tf32
    # (3) this path
    y = net(xg)
    B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y, lg.unsqueeze(1)).backward()
    taps = net.forward_debug(xg)
    table, bad = [], []
    for k in taps32:
        if k not in taps:
            continue
        e_o, e_t = rel(taps[k], taps32[k]), rel(taps16[k], taps32[k])
        table.append(("out:" + k, e_o, e_t))
        if e_o > CAL_FACTOR * e_t + 2e-3:
            bad.append(table[-1])
    e_o, e_t = rel(y, y32.detach()), rel(y16, y32.detach())
    table.append(("logits", e_o, e_t))
    if e_o > CAL_FACTOR * e_t + 2e-3:
        bad.append(table[-1])
    names = [n for n, _ in net.named_parameters()]
    slopes = []
    for n, p in net.named_parameters():
        if n.endswith("conv.bias") and (n[:-len("conv.bias")] + "act.weight") in names:
            continue  # dead bias: true gradient 0, both references hold rounding noise
        if n.endswith("act.weight"):
            slopes.append((n, p.grad.item(), g16[n].item(), g32[n].item()))
            continue
        e_o, e_t = rel(p.grad, g32[n]), rel(g16[n], g32[n])
        table.append(("grad:" + n, e_o, e_t))
        if e_o > CAL_FACTOR * e_t + 1e-2:
            bad.append(table[-1])
    # PReLU slopes: ONE scalar per layer, a full-tensor sum with heavy cancellation -- relative error per scalar is
    # ill-conditioned (stock torch bf16 itself is 2-30 % off on single slopes, and it keeps InstanceNorm/PReLU
    # activations in fp32 under autocast while this path stores them in bf16).  They are calibrated as ONE vector
    # (all layers), and each scalar against the vector's largest entry.
    so, st_, s32 = (torch.tensor([r[k] for r in slopes], dtype=torch.float64) for k in (1, 2, 3))
    e_o, e_t = rel(so, s32), rel(st_, s32)
    table.append(("grad:act.weight (all %d slopes as one vector)" % len(slopes), e_o, e_t))
    if e_o > CAL_FACTOR * e_t + 1e-2:
        bad.append(table[-1])
    scale = float(s32.abs().max())
    for n, a, b, c in slopes:
        table.append(("slope:" + n, abs(a - c) / scale, abs(b - c) / scale))
        if abs(a - c) > CAL_FACTOR * abs(b - c) + 0.1 * scale:
            bad.append(table[-1])
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"bf16_calibration_{case}.json"), "w") as f:
        json.dump({"case": case, "columns": ["tensor", "err_b200seg_vs_fp32", "err_torch_bf16_vs_fp32"],
                   "rows": table}, f, indent=0)
    outs = [r for r in table if r[0].startswith("out:")]
    grs = [r for r in table if r[0].startswith("grad:")]
    print(f"\n[{case}] layer outputs: ours max {max(r[1] for r in outs):.3e} / torch-bf16 max {max(r[2] for r in outs):.3e}; "
          f"logits {table[len(outs)][1]:.3e} / {table[len(outs)][2]:.3e}; "
          f"gradients: ours max {max(r[1] for r in grs):.3e} / torch-bf16 max {max(r[2] for r in grs):.3e}")
    assert not bad, f"drifts more than {CAL_FACTOR} x torch bf16: {bad}"
