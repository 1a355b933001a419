"""Host-side logic of the UNet engine on CPU: the forward/reverse plans (zero-copy skip
concatenation, in-place gradient fan-in, fused residual addends) are run with the kernels
replaced by a torch emulation (tests/_torch_ops.py) and compared with autograd on the oracle."""
import pytest
import torch

import ct_image_segmentation_b200.unet as U
from oracle import monai_ref as O

from . import _torch_ops

CASES = [
    (3, 1, [4, 8, 8, 16, 16], [2, 2, 2, 2], 2, (1, 1, 32, 48, 32)),
    (3, 2, [4, 6, 8], [2, 2], 1, (2, 2, 8, 8, 12)),
    (3, 1, [4, 8, 8], [2, 1], 0, (1, 1, 8, 8, 8)),
    (2, 3, [4, 8, 8, 16, 16], [2, 2, 2, 2], 2, (2, 3, 32, 48)),
    (2, 1, [4, 8, 8, 16, 16], [2, 2, 2, 2], 0, (1, 1, 32, 32)),
]


@pytest.mark.parametrize("dims,inc,ch,st,res,shape", CASES)
def test_plan_matches_autograd(monkeypatch, dims, inc, ch, st, res, shape):
    monkeypatch.setattr(U, "ops", _torch_ops)
    torch.manual_seed(12342)
    ref = O.UNet(dims, inc, 10, ch, st, num_res_units=res)
    net = U.UNet(dims, inc, 10, ch, st, num_res_units=res, dtype=torch.float32)
    net.load_state_dict(ref.state_dict())
    x = torch.randn(*shape, requires_grad=True)
    y_ref = ref(x)
    g = torch.randn_like(y_ref)
    y_ref.backward(g)

    saved = {}
    out = net._run_forward(_torch_ops.to_channels_last(x.detach(), torch.float32), saved)
    y = _torch_ops.from_channels_last(out, dims)
    torch.testing.assert_close(y, y_ref.detach(), rtol=1e-3, atol=5e-4)
    grads, gx = net._run_backward(saved, _torch_ops.to_channels_last(g, torch.float32), True)
    assert not saved  # every saved activation was consumed
    torch.testing.assert_close(_torch_ops.from_channels_last(gx, dims), x.grad, rtol=5e-3, atol=5e-4)
    ref_params = dict(ref.named_parameters())
    for name, p in net.named_parameters():
        assert p in grads, name
        rg = ref_params[name].grad
        scale = max(rg.abs().max().item(), 1e-3)
        if name.endswith("conv.bias") and name[:-len("conv.bias")] + "act.weight" in ref_params:
            # bias in front of an affine-less InstanceNorm: true gradient is 0, both sides hold
            # rounding noise -> absolute bound against the layer's weight grad
            wscale = ref_params[name[:-4] + "weight"].grad.abs().max().item()
            assert (grads[p] - rg).abs().max().item() <= 1e-2 * wscale + 1e-3, name
            continue
        assert (grads[p] - rg).abs().max().item() <= 5e-3 * scale + 1e-4, name


def test_gradient_sink_and_level_hook(monkeypatch):
    """With a gradient sink the reverse plan writes every live gradient into the bucket views (dead biases
    stay zero, nothing is handed to autograd), and the level hook fires once per depth, deepest first, at a
    point where the gradients of that level and everything below it are already in the bucket."""
    from ct_image_segmentation_b200.parallel import GradientBucket
    monkeypatch.setattr(U, "ops", _torch_ops)
    torch.manual_seed(5)
    dims, inc, ch, st, res, shape = CASES[0]
    net = U.UNet(dims, inc, 10, ch, st, num_res_units=res, dtype=torch.float32)
    x = torch.randn(*shape)
    g = torch.randn(shape[0], 10, *shape[2:])

    saved = {}
    net._run_forward(_torch_ops.to_channels_last(x, torch.float32), saved)
    grads, _ = net._run_backward(saved, _torch_ops.to_channels_last(g, torch.float32), False)

    bucket = GradientBucket(net.parameters())
    sink = dict(zip(bucket.params, bucket.views))
    names = {id(p): n for n, p in net.named_parameters()}
    fired = []

    def hook(depth):
        # complete at hook time: everything but the down layers of the levels above `depth` -- the first
        # parameters of the depth-first bucket -- i.e. one contiguous range up to the END of the bucket
        outer, m = [], net.model
        for _ in range(depth):
            outer += list(m[0].parameters())
            m = m[1].submodule
        assert all(a is b for a, b in zip(bucket.params, outer))
        for p in bucket.params[len(outer):]:
            torch.testing.assert_close(sink[p], grads[p].view_as(p), rtol=0, atol=0, msg=names[id(p)])
        fired.append(depth)

    net.bind_grad_sink(sink, hook)
    saved = {}
    net._run_forward(_torch_ops.to_channels_last(x, torch.float32), saved)
    grads2, _ = net._run_backward(saved, _torch_ops.to_channels_last(g, torch.float32), False)
    net.bind_grad_sink(None)
    assert fired == [4, 3, 2, 1]
    assert not any(p in grads2 for p in net.parameters())
    for p in net.parameters():
        torch.testing.assert_close(sink[p], grads[p].view_as(p), rtol=0, atol=0, msg=names[id(p)])
