"""End-to-end parity of the B200 UNet + Dice path against the CPU oracle -- `pytest -m gpu`.

Same seeded inputs and weights on both sides: per-layer outputs, logits,
Dice loss, every parameter gradient, argmax label maps.  fp32 check mode: 1e-4 relative, label map
bit-exact; bf16: 1e-2 relative, Dice within 1e-3.
"""
import pytest
import torch

import ct_image_segmentation_b200 as B
from ct_image_segmentation_b200.capstone.training.base_trainer import BaseUNet2D
from ct_image_segmentation_b200.capstone.volumetric.base_trainer import BaseUNet3D
from oracle import monai_ref as O

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def make_pair(dims, inc, channels, strides, res, dtype, seed=12342):
    torch.manual_seed(seed)
    ref = O.UNet(dims, inc, 10, channels, strides, num_res_units=res)
    # move PReLU slopes and biases off their init so their gradients are exercised
    with torch.no_grad():
        for n, p in ref.named_parameters():
            if n.endswith("act.weight"):
                p.uniform_(0.1, 0.4)
    net = B.UNet(dims, inc, 10, channels, strides, num_res_units=res, dtype=dtype)
    net.load_state_dict(ref.state_dict())
    return ref, net.to(DEV)


def sparse_labels(n, sp, seed=0):
    g = torch.Generator().manual_seed(seed)
    lab = torch.zeros(n, *sp, dtype=torch.int64)
    for i in range(n):
        for c in range(1, 10):
            lo = [int(torch.randint(0, max(1, s - s // 3), (1,), generator=g)) for s in sp]
            sl = tuple(slice(l, l + max(2, s // 4)) for l, s in zip(lo, sp))
            lab[(i, *sl)] = c
    return lab


def dead_bias(name, ref_params):
    return name.endswith("conv.bias") and (name[:-len("conv.bias")] + "act.weight") in ref_params


CASES = [
    # dims, in, channels, strides, res, input shape
    (3, 1, [16, 32, 64, 128, 256], [2, 2, 2, 2], 2, (2, 1, 32, 48, 32)),
    (3, 1, [8, 16, 16], [2, 2], 0, (1, 1, 16, 16, 24)),
    (3, 2, [8, 12, 16, 24], [2, 1, 2], 1, (1, 2, 16, 16, 16)),
    (2, 3, [16, 32, 32, 64, 64], [2, 2, 2, 2], 2, (2, 3, 64, 96)),
    (2, 1, [8, 16, 32, 32, 64], [2, 2, 2, 2], 0, (1, 1, 64, 64)),
]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims,inc,ch,st,res,shape", CASES)
def test_unet_dice_parity(dims, inc, ch, st, res, shape, dtype):
    ref, net = make_pair(dims, inc, ch, st, res, dtype)
    torch.manual_seed(1)
    x = torch.randn(*shape)
    lab = sparse_labels(shape[0], shape[2:])
    fp32 = dtype == torch.float32
    tol = 1e-4 if fp32 else 1e-2

    # ---- reference (CPU oracle) with per-layer taps
    taps_ref = {}
    hooks = []
    for name, m in ref.named_modules():
        if isinstance(m, O.Convolution):
            hooks.append(m.register_forward_hook(lambda mod, i, o, name=name: taps_ref.__setitem__(name, o.detach())))
            hooks.append(m.conv.register_forward_hook(
                lambda mod, i, o, name=name + ".conv": taps_ref.__setitem__(name, o.detach())))
    y_ref = ref(x)
    for h in hooks:
        h.remove()
    loss_ref = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y_ref, lab.unsqueeze(1))
    loss_ref.backward()

    # ---- B200 path
    xd = x.to(DEV)
    y = net(xd)
    assert tuple(y.shape) == tuple(y_ref.shape) and y.dtype == dtype
    loss = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y, lab.to(DEV).unsqueeze(1))
    loss.backward()

    taps = net.forward_debug(xd)
    worst = max((rel(taps[k], v), k) for k, v in taps_ref.items() if k in taps)
    n_conv = sum(isinstance(m, O.Convolution) for m in ref.modules())
    assert sum(k.endswith(".conv") and k in taps for k in taps_ref) >= n_conv - 1  # every conv output compared
    # bf16, END-TO-END drift (every layer's input already carries the rounding of all layers before it): stock
    # torch bf16 drifts 1.3e-2 on the same nets (tests/test_gpu_parity_shapes.py calibration, profiles/
    # r2_bf16_calibration_*.json: ours 1.1e-2 / 8e-3 logits vs torch 1.3e-2 / 9.6e-3).  The north-star 1e-2 bound is
    # asserted per layer on identical inputs (layer-local tests: worst 1.7e-3).
    assert worst[0] < (tol if fp32 else 2e-2), f"per-layer output {worst}"
    assert rel(y, y_ref.detach()) < (tol if fp32 else 2e-2)
    assert abs(loss.item() - loss_ref.item()) < (1e-5 if fp32 else 1e-3)   # north_star: Dice within 1e-3
    if not fp32:
        # bf16 gradients: layer by layer on identical inputs at 1e-2 (test_unet_layerwise_backward and, at the
        # benchmarked shapes, tests/test_gpu_parity_shapes.py); end to end they are calibrated against stock torch
        # bf16 (test_bf16_drift_calibrated_against_torch_bf16) and checked by direction below.
        return

    ref_params = dict(ref.named_parameters())
    bad = []
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        rg = ref_params[name].grad
        if dead_bias(name, ref_params):
            wn = ref_params[name[:-4] + "weight"].grad.abs().max().item()
            if (p.grad.cpu() - rg).abs().max().item() > (1e-3 if fp32 else 5e-2) * wn + 1e-6:
                bad.append((name, "dead-bias abs"))
            continue
        e = rel(p.grad, rg)
        # Whole-network gradients cross PReLU kinks: a 1e-6 difference in a pre-activation that
        # sits at zero flips its slope, and with sparse labels a single high-gradient voxel can carry
        # 1e-3 of the gradient norm.  The 1e-4 per-layer bound is asserted layer-locally (identical
        # inputs) in test_unet_layerwise_backward; here the bound is the kink-tolerant 5e-3.
        if e >= (5e-2 if name.endswith("act.weight") else 5e-3):
            bad.append((name, e))
    assert not bad, f"parameter gradients out of tolerance: {bad}"

    if fp32:  # label map bit-exact in check mode (away from ties created by 1e-6 logit noise)
        lm_ref = O.squash_predictions(y_ref.detach())
        lm = B.squash_predictions(y.detach())
        mism = (lm.cpu() != lm_ref)
        if mism.any():
            top2 = torch.softmax(y_ref.detach(), 1).topk(2, dim=1).values
            gap = (top2[:, 0] - top2[:, 1])[mism]
            assert float(gap.max()) < 1e-5, f"{int(mism.sum())} label mismatches with margin up to {float(gap.max())}"


def cosine(a, b):
    a, b = a.double().cpu().reshape(-1), b.double().cpu().reshape(-1)
    return (a @ b / (a.norm() * b.norm()).clamp_min(1e-300)).item()


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("dims,inc,ch,st,res,shape", CASES[:4])
def test_unet_layerwise_backward(dims, inc, ch, st, res, shape, dtype):
    """Gradients layer by layer with IDENTICAL inputs (north_star: per-layer gradients within 1e-4
    in fp32 check mode, 1e-2 in bf16).  A whole-network comparison of bf16 gradients is not meaningful at 1e-2: bf16
    storage noise (2^-8) saturates after a few layers whatever the summation order, flips PReLU
    masks of near-zero activations and the gradient is discontinuous there ("bf16
    parity").  So each Convolution's backward is checked in situ: its own inputs (x, c, mean, rstd,
    incoming gradient) are taken from the GPU run and its outputs recomputed with torch fp32."""
    import torch.nn.functional as F
    from ct_image_segmentation_b200 import ops
    from ct_image_segmentation_b200.unet import Convolution
    ref, net = make_pair(dims, inc, ch, st, res, dtype)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    torch.manual_seed(1)
    x = torch.randn(*shape)
    lab = sparse_labels(shape[0], shape[2:])
    saved = {}
    # use_cols=False: the taps must hold every layer's real input (the im2col first-layer path is
    # covered by test_first_layer_im2col and by the whole-network tests)
    out = net._run_forward(ops.to_channels_last(x.to(DEV), dtype), saved, use_cols=False)
    lg = ops.from_channels_last(out, dims).detach().requires_grad_(True)
    loss = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(lg, lab.to(DEV).unsqueeze(1))
    loss.backward()
    taps = {}
    grads, _ = net._run_backward(saved, ops.to_channels_last(lg.grad, dtype), False, taps)
    names = {m: n for n, m in net.named_modules()}
    taps = {m: t for m, t in taps.items() if isinstance(m, Convolution)}  # (residual-unit taps: test_gpu_parity_shapes)
    assert len(taps) == sum(isinstance(m, Convolution) for m in net.modules())
    bad = []
    for m, t in taps.items():
        name, g = names[m], m.geom
        g_out, g_c = t["g_out"].float().cpu(), t["g_c"].float().cpu()
        if t["c"] is not None:
            n, c = g_c.shape[0], g_c.shape[-1]
            # statistics of zero-padded 10-class tensors are kept for the padded channel count
            mean = t["mean"].cpu().view(n, -1)[:, :c].reshape(n, 1, 1, 1, c)
            rstd = t["rstd"].cpu().view(n, -1)[:, :c].reshape(n, 1, 1, 1, c)
            h = (t["c"].float().cpu() - mean) * rstd
            alpha = m.act.weight.detach().cpu()
            gt = torch.where(h > 0, g_out, alpha * g_out)
            ref_gc = rstd * (gt - gt.mean(dim=(1, 2, 3), keepdim=True) - h * (gt * h).mean(dim=(1, 2, 3), keepdim=True))
            e = rel(g_c, ref_gc)
            if e >= tol:
                bad.append((name, "in+prelu bwd", e))
            terms = torch.where(h > 0, torch.zeros_like(h), g_out * h)
            da = grads[m.act.weight].item()
            if abs(da - terms.double().sum().item()) > 1e-3 * terms.abs().double().sum().item() + 1e-12:
                bad.append((name, "dalpha", da, terms.double().sum().item()))
        # wgrad / bias grad from the SAME (x, g_c)
        xin = t["x"].float().cpu().permute(0, 4, 1, 2, 3)
        gy = g_c.permute(0, 4, 1, 2, 3)
        if dims == 2:
            xin, gy = xin.squeeze(2), gy.squeeze(2)
        w = torch.zeros_like(m.conv.weight.detach().cpu()).requires_grad_(True)
        p = (g.kernel - 1) // 2
        if g.transposed:
            f = F.conv_transpose2d if dims == 2 else F.conv_transpose3d
            y = f(xin, w, stride=g.stride, padding=p, output_padding=g.stride - 1)
        else:
            f = F.conv2d if dims == 2 else F.conv3d
            y = f(xin, w, stride=g.stride, padding=p)
        (gw_ref,) = torch.autograd.grad(y, w, gy)
        e = rel(grads[m.conv.weight], gw_ref)
        if e >= tol:
            bad.append((name, "wgrad", e))
        gb_ref = gy.sum(dim=[0] + list(range(2, gy.dim())))
        if (grads[m.conv.bias].cpu() - gb_ref).abs().max().item() > 1e-2 * gy.abs().sum(dim=[0] + list(range(2, gy.dim()))).max().item():
            bad.append((name, "bias grad"))
    assert not bad, f"layer-local backward out of tolerance: {bad}"


@pytest.mark.parametrize("dims,inc,ch,st,res,shape", CASES[:1] + CASES[3:4])
def test_unet_bf16_gradient_direction_vs_fp32_oracle(dims, inc, ch, st, res, shape):
    """Whole-network bf16 gradients against the fp32 oracle: same direction (cosine), Dice within
    1e-3 -- the bound that is meaningful across PReLU mask flips (see the test above)."""
    ref, net = make_pair(dims, inc, ch, st, res, torch.bfloat16)
    torch.manual_seed(1)
    x = torch.randn(*shape)
    lab = sparse_labels(shape[0], shape[2:])
    loss_ref = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(ref(x), lab.unsqueeze(1))
    loss_ref.backward()
    loss = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(net(x.to(DEV)), lab.to(DEV).unsqueeze(1))
    loss.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-3
    ref_params = dict(ref.named_parameters())
    cos = {n: cosine(p.grad, ref_params[n].grad) for n, p in net.named_parameters()
           if n.endswith("weight") and not n.endswith("act.weight")}
    worst = min(cos.items(), key=lambda kv: kv[1])
    assert worst[1] > 0.9, f"bf16 weight-gradient direction off: {worst}"


def test_state_dict_roundtrip_and_repack():
    ref, net = make_pair(3, 1, [8, 16, 16], [2, 2], 2, torch.float32)
    assert list(net.state_dict().keys()) == list(ref.state_dict().keys())
    x = torch.randn(1, 1, 8, 8, 8, device=DEV)
    y0 = net(x).detach().clone()
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.5)  # in-place update, as an optimiser does -> packed weights must refresh
    y1 = net(x).detach()
    assert rel(y1, y0) > 1e-3
    with torch.no_grad():
        for p, q_ in zip(ref.parameters(), net.parameters()):
            p.copy_(q_.cpu())
    assert rel(y1, ref(x.cpu()).detach()) < 1e-4


def test_input_gradient_and_eval_mode():
    ref, net = make_pair(3, 1, [8, 16, 16], [2, 2], 2, torch.float32)
    x = torch.randn(1, 1, 16, 16, 16)
    xr = x.clone().requires_grad_(True)
    ref(xr).square().sum().backward()
    xd = x.to(DEV).requires_grad_(True)
    net.eval()
    net(xd).float().square().sum().backward()
    assert rel(xd.grad, xr.grad) < 1e-4
    with torch.no_grad():
        assert rel(net(x.to(DEV)), ref(x).detach()) < 1e-4
    with pytest.raises(ValueError):
        net(torch.zeros(1, 1, 10, 16, 16, device=DEV))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 1, 16, 16, 16))


@pytest.mark.parametrize("exclude_missing", [False, True])
def test_lightning_module_3d_step(exclude_missing):
    torch.manual_seed(12342)
    mod = BaseUNet3D(filters=[8, 16, 16, 32, 32], use_res_units=True, loss_fx=["Dice"],
                     exclude_missing=exclude_missing, dtype=torch.float32).to(DEV)
    ref = O.UNet(3, 1, 10, [8, 16, 16, 32, 32], [2, 2, 2, 2], num_res_units=2)
    ref.load_state_dict(mod.unet.state_dict())
    n, sp = 2, (16, 32, 16)
    images = torch.randn(n, 1, *sp)
    lab = sparse_labels(n, sp, seed=3)
    masks = torch.stack([(lab == c) for c in range(1, 10)], dim=1).to(torch.uint8)
    ind = torch.ones(n, 9)
    if exclude_missing:
        ind[1, 3] = 0
        ind[0, 7] = 0
    batch = (images.to(DEV), masks.to(DEV), ind.to(DEV))
    loss = mod.training_step(batch, 0)
    loss.backward()
    # oracle
    y_ref = ref(images)
    d = O.MultipleLossWrapper(["Dice"], exclude_missing)(y_ref, O.squash_masks(masks), ind)
    assert abs(loss.item() - d["Dice"].item()) < 1e-5
    dm, dpc = O.dice_metric(O.squash_predictions(y_ref.detach()), O.squash_masks(masks))
    assert abs(mod.logged["Mean Dice Score (train)"].item() - dm.item()) < 1e-6
    assert "BrainStem Dice (train)" in mod.logged and "Dice Loss (train)" in mod.logged
    opt = mod.configure_optimizers()
    opt.step()
    mod.validation_step(batch, 0)
    assert "Mean Dice Score (val)" in mod.logged


def test_lightning_module_2d_step():
    torch.manual_seed(1)
    mod = BaseUNet2D(filters=[8, 16, 16, 32, 32], use_res_units=True, loss_fx=["Dice"],
                     dtype=torch.bfloat16).to(DEV)
    images = torch.randn(2, 1, 64, 64, device=DEV)
    masks = (torch.rand(2, 9, 64, 64, device=DEV) > 0.9).to(torch.uint8)
    loss = mod.training_step((images, masks, torch.ones(2, 9, device=DEV)), 0)
    loss.backward()
    assert torch.isfinite(loss) and all(p.grad is not None for p in mod.unet.parameters())


def test_lightning_module_2d_downsample_and_boundary_vs_oracle():
    """BaseUNet2D as the reference's default recipe uses it (capstone/training/base_trainer.py:22-118): 3 windowed
    channels -> `conv1x1` (downsample=True) -> UNet, losses Boundary + Dice + Focal with the batch's distance maps
    (4-tuple batch), Dice metric -- against the oracle with the same weights: total loss, every logged loss, the
    metric, and the gradient that reaches the 1x1 convolution through the network's input gradient."""
    torch.manual_seed(4)
    mod = BaseUNet2D(filters=[8, 16, 16, 32, 32], use_res_units=True, downsample=True,
                     loss_fx=["Focal", "Dice", "Boundary"], transform_degree=1, dtype=torch.float32).to(DEV)
    assert mod.unet.in_channels == 1 and list(mod.loss_func.losses) == ["Boundary", "Dice", "Focal"]  # sorted
    ref = O.UNet(2, 1, 10, [8, 16, 16, 32, 32], [2, 2, 2, 2], num_res_units=2)
    ref.load_state_dict(mod.unet.state_dict())
    c11 = torch.nn.Conv2d(3, 1, 1)
    c11.load_state_dict(mod.conv1x1.state_dict())
    n, sp = 2, (64, 48)
    images = torch.randn(n, 3, *sp)
    lab = sparse_labels(n, sp, seed=5)
    masks = torch.stack([(lab == c) for c in range(1, 10)], dim=1).to(torch.uint8)
    dist = torch.randn(n, 9, *sp) * 0.02
    ind = torch.ones(n, 9)
    loss = mod.training_step((images.to(DEV), masks.to(DEV), ind.to(DEV), dist.to(DEV)), 0)
    loss.backward()
    y_ref = ref(c11(images))
    want = O.MultipleLossWrapper(["Boundary", "Dice", "Focal"])(y_ref, O.squash_masks(masks), ind, dist)
    total = torch.stack(list(want.values())).sum()
    total.backward()
    assert abs(loss.item() - total.item()) < 1e-5 * max(1.0, abs(total.item()))
    for k, v in want.items():
        assert abs(mod.logged[f"{k} Loss (train)"].item() - v.item()) < 1e-5, k
    dm, _ = O.dice_metric(O.squash_predictions(y_ref.detach()), O.squash_masks(masks))
    assert abs(mod.logged["Mean Dice Score (train)"].item() - dm.item()) < 1e-6
    assert rel(mod.conv1x1.weight.grad, c11.weight.grad) < 5e-3 and rel(mod.conv1x1.bias.grad, c11.bias.grad) < 5e-3
    assert BaseUNet2D.add_model_specific_args(__import__("argparse").ArgumentParser()).parse_args([]).batch_size == 128


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_wgrad_side_stream_is_bitwise_identical(dtype):
    """Weight gradients launched on the second stream (parallel graph branch) == single-stream backward."""
    ref, net = make_pair(3, 1, [16, 32, 64], [2, 2], 2, dtype)
    torch.manual_seed(3)
    x = torch.randn(2, 1, 32, 32, 32, device=DEV)
    lab = torch.randint(0, 10, (2, 1, 32, 32, 32), device=DEV)
    fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    grads = []
    for side in (False, True, True):
        net.enable_wgrad_stream(side)
        for p in net.parameters():
            p.grad = None
        fx(net(x), lab).backward()
        torch.cuda.synchronize()
        grads.append([p.grad.clone() for p in net.parameters()])
    net.enable_wgrad_stream(False)
    for a, b, c in zip(*grads):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_flat_adam_matches_torch_adam():
    """FlatAdam (one kernel on the flat parameter / gradient buffers) == torch.optim.Adam over several steps."""
    torch.manual_seed(0)
    shapes = [(16, 1, 3, 3, 3), (16,), (1,), (32, 16, 3, 3, 3), (7,)]
    pa = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa = torch.optim.Adam(pa, lr=3e-3)
    ob = B.FlatAdam(pb, lr=3e-3)
    for step in range(6):
        for x, y in zip(pa, pb):
            g = torch.randn_like(x) * (10.0 ** (step - 3))
            x.grad, y.grad = g, g.clone()
        oa.step()
        ob.step()
    for x, y in zip(pa, pb):
        assert rel(y.detach(), x.detach()) < 1e-6
    assert all(y.data_ptr() >= ob.flat.data_ptr() for y in pb)  # parameters live in the flat buffer


def test_flat_adam_state_dict_roundtrip():
    """the flat moments and the step count survive state_dict()/load_state_dict(), in torch.optim.Adam's
    own layout (a checkpoint resumes under either optimiser); unsupported options raise instead of being ignored."""
    torch.manual_seed(0)
    shapes = [(8, 2, 3, 3), (8,), (1,)]
    pa = [torch.nn.Parameter(torch.randn(s, device=DEV)) for s in shapes]
    pb = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    pc = [torch.nn.Parameter(p.detach().clone()) for p in pa]
    oa, ob = torch.optim.Adam(pa, lr=1e-2), B.FlatAdam(pb, lr=1e-2)
    gs = [[torch.randn(s, device=DEV) for s in shapes] for _ in range(5)]
    for k in range(3):
        for x, y, g in zip(pa, pb, gs[k]):
            x.grad, y.grad = g.clone(), g.clone()
        oa.step()
        ob.step()
    sd = ob.state_dict()
    assert len(sd["state"]) == 3 and float(sd["state"][0]["step"]) == 3.0
    # resume: FlatAdam <- FlatAdam state, torch Adam <- FlatAdam state
    with torch.no_grad():
        for y, z in zip(pb, pc):
            z.copy_(y)
    oc = B.FlatAdam(pc, lr=1.0)
    oc.load_state_dict(sd)
    assert oc.steps == 3 and oc.param_groups[0]["lr"] == 1e-2
    od = torch.optim.Adam([torch.nn.Parameter(y.detach().clone()) for y in pb], lr=1e-2)
    od.load_state_dict(sd)
    pd_ = od.param_groups[0]["params"]
    for k in range(3, 5):
        for x, y, z, w, g in zip(pa, pb, pc, pd_, gs[k]):
            x.grad, y.grad, z.grad, w.grad = g.clone(), g.clone(), g.clone(), g.clone()
        for o in (oa, ob, oc, od):
            o.step()
    for x, y, z, w in zip(pa, pb, pc, pd_):
        assert rel(y.detach(), x.detach()) < 1e-6 and torch.equal(y.detach(), z.detach())
        assert rel(w.detach(), x.detach()) < 1e-6
    with pytest.raises(NotImplementedError):
        B.FlatAdam(pb, weight_decay=0.1)
    with pytest.raises(NotImplementedError):
        ob.add_param_group({"params": [torch.nn.Parameter(torch.zeros(1, device=DEV))]})


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_flat_adam_eager_steps_refresh_packed_weights(dtype):
    """FlatAdam writes the parameters through a raw pointer; the packed-weight caches key on
    ``_version``, so every step must bump it.  Eager (no graph) training with FlatAdam must follow torch.optim.Adam
    on the oracle: the loss changes from step to step and the parameters / logits stay together."""
    ref, net = make_pair(3, 1, [8, 16, 16], [2, 2], 2, dtype)
    torch.manual_seed(5)
    x = torch.randn(1, 1, 16, 16, 16)
    lab = torch.randint(0, 10, (1, 1, 16, 16, 16))
    fr = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    oa = torch.optim.Adam(ref.parameters(), lr=1e-2)
    ob = B.FlatAdam(net.parameters(), lr=1e-2)
    v0 = [p._version for p in net.parameters()]
    losses, losses_ref = [], []
    for _ in range(4):
        oa.zero_grad()
        lr_ = fr(ref(x), lab)
        lr_.backward()
        oa.step()
        losses_ref.append(lr_.item())
        for p in net.parameters():
            p.grad = None
        lb = fx(net(x.to(DEV)), lab.to(DEV))
        lb.backward()
        ob.step()
        losses.append(lb.item())
    assert all(p._version > v for p, v in zip(net.parameters(), v0))
    assert len({round(l, 6) for l in losses}) == 4, losses  # stale packed weights would repeat the first loss
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    for a, b in zip(losses, losses_ref):
        assert abs(a - b) < tol, (losses, losses_ref)
    with torch.no_grad():
        y, y_ref = net(x.to(DEV)), ref(x)
    # (Adam's first steps are sign-like: bf16 gradient noise moves individual weights by +-lr, so the bf16 logits
    # only have to stay in the oracle's neighbourhood; stale weights would leave them at the INITIAL network's)
    assert rel(y, y_ref) < (5e-3 if dtype == torch.float32 else 0.3)


@pytest.mark.parametrize("use_graph", [False, True])
def test_graphed_train_step_matches_autograd(use_graph):
    """GraphedTrainStep (weight gradients written straight into the flat bucket, CUDA-graph replay) gives the
    gradients and the loss of the plain autograd call, bit for bit, and keeps doing so over replays."""
    ref, net = make_pair(3, 1, [16, 32, 64, 64], [2, 2, 2], 2, torch.bfloat16)
    torch.manual_seed(11)
    x = torch.randn(2, 1, 32, 32, 32, device=DEV)
    lab = torch.randint(0, 10, (2, 32, 32, 32), device=DEV)
    fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    loss0 = fx(net(x), lab.unsqueeze(1))
    loss0.backward()
    # drop the autograd graph: its AccumulateGrad nodes are bound to the default stream and would be reused
    # (and synchronised with) inside the capture
    loss0 = loss0.detach()
    want = [p.grad.clone() for p in net.parameters()]
    for p in net.parameters():
        p.grad = None
    train = B.GraphedTrainStep(net, fx, None, x, lab, use_graph=use_graph)
    names = [n for n, _ in net.named_parameters()]
    for _ in range(3):
        for v, w in zip(train.bucket.views, want):
            if w.any():
                v.fill_(float("nan"))  # every live gradient is rewritten by the step
        loss = train(None, None)
        torch.cuda.synchronize()
        assert torch.equal(loss, loss0.detach())
        for n, p, v, w in zip(names, train.bucket.params, train.bucket.views, want):
            assert p.grad.data_ptr() == v.data_ptr()
            assert torch.equal(v, w), n  # dead biases: exactly zero on both sides (never written)
    net.enable_wgrad_stream(False)
