"""A/B of the fused dgrad + InstanceNorm-backward sums at the cfg3 shapes (us per call, CUDA-graph replays)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import _graph_time_us
from ct_image_segmentation_b200 import _lib, ops

DEV, DT = torch.device("cuda", 0), torch.bfloat16
torch.cuda.set_device(0)
for (c, n, p) in ((10, 2, 128), (16, 2, 64)):
    g = ops.ConvGeom(3, c, c, 3, 1, False)
    sp = (p, p, p)
    def act():
        t = ops.alloc_activation(n, sp, c, DT, DEV); t.copy_(torch.randn(t.shape, device=DEV)); return t
    cprev, dy, res = act(), act(), act()
    dx, gc = ops.alloc_like(cprev), ops.alloc_like(cprev)
    w = torch.randn(c, c, 3, 3, 3, device=DEV) * 0.1
    wd = ops.pack_weight(g, _lib.W_CONV_DGRAD, w, DT)
    mean, rstd = ops.instnorm_stats(cprev)
    a = torch.full((1,), 0.25, device=DEV)
    t1 = _graph_time_us(lambda: ops.conv_dgrad(g, dy, wd, dx, residual=res))
    t2 = _graph_time_us(lambda: ops.instnorm_prelu_bwd(cprev, mean, rstd, a, dx, gc))
    h = ops.conv_dgrad_instnorm_partials(g, dy, wd, dx, cprev, mean, rstd, a, residual=res)
    t3 = _graph_time_us(lambda: ops.conv_dgrad_instnorm_partials(g, dy, wd, dx, cprev, mean, rstd, a, residual=res))
    t4 = _graph_time_us(lambda: ops.instnorm_prelu_bwd_from_partials(cprev, mean, rstd, a, dx, gc, h))
    print(f"{c}ch @{p}^3 x{n}: dgrad {t1:.1f} + IN bwd {t2:.1f} = {t1 + t2:.1f} us   |  fused dgrad {t3:.1f} + IN bwd rest {t4:.1f} = {t3 + t4:.1f} us")
