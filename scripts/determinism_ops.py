"""Run-to-run determinism of single kernels at the cfg5 full-resolution shapes (batch x 160^3).
python scripts/determinism_ops.py [batch] [patch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ct_image_segmentation_b200 import _lib, ops  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4
P = int(sys.argv[2]) if len(sys.argv) > 2 else 160
DEV, DT = torch.device("cuda", 0), torch.bfloat16
lib = _lib.load()


def act(n, sp, c):
    t = ops.alloc_activation(n, sp, c, DT, DEV)
    t.copy_(torch.randn(t.shape, device=DEV))
    return t


def check(name, fn, reps=4):
    outs = []
    for _ in range(reps):
        o = fn()
        torch.cuda.synchronize()
        flat = []
        for t in (o if isinstance(o, (tuple, list)) else [o]):
            flat += list(t) if isinstance(t, (tuple, list)) else [t]
        outs.append([t.clone() for t in flat if t is not None])
    ok = all(all(torch.equal(a, b) for a, b in zip(outs[0], o)) for o in outs[1:])
    kern = lib.b200seg_last_launch().decode()
    extra = ""
    if not ok:
        d = max(max((a.float() - b.float()).abs().max().item() for a, b in zip(outs[0], o)) for o in outs[1:])
        extra = f" max abs diff {d:.3e}"
    print(f"{'OK  ' if ok else 'DIFF'} {name} [{kern}]{extra}", flush=True)


torch.manual_seed(0)
z = act(N, (P, P, P), 10)
lab = torch.randint(0, 10, (N, P, P, P), device=DEV, dtype=torch.uint8)
gi, gp = torch.rand(N, 10, device=DEV), torch.rand(N, 10, device=DEV)
dz = ops.alloc_like(z)
check("dice fwd+metric", lambda: ops.softmax_dice_metric_sums(z, lab))
check("dice bwd", lambda: ops.softmax_dice_bwd(z, lab, gi, gp, dlogits=dz))
g = ops.ConvGeom(3, 10, 10, 3, 1, False)
x, dy = act(N, (P, P, P), 10), act(N, (P, P, P), 10)
w = torch.randn(10, 10, 3, 3, 3, device=DEV) * 0.05
wf, wd = ops.pack_weight(g, _lib.W_CONV_FPROP, w, DT), ops.pack_weight(g, _lib.W_CONV_DGRAD, w, DT)
y, dx = ops.alloc_like(x), ops.alloc_like(x)
b = torch.zeros(10, device=DEV)
check("head fprop", lambda: ops.conv_fprop(g, x, wf, b, y))
check("head dgrad", lambda: ops.conv_dgrad(g, dy, wd, dx))
check("head wgrad", lambda: ops.conv_wgrad(g, x, dy, want_bias=True))
mean, rstd = ops.instnorm_stats(x)
alpha = torch.full((1,), 0.25, device=DEV)
check("IN stats 10ch", lambda: ops.instnorm_stats(x))
check("IN bwd 10ch", lambda: (ops.instnorm_prelu_bwd(x, mean, rstd, alpha, dy, dx), dx))
gt = ops.ConvGeom(3, 64, 10, 3, 2, True)
xt = act(N, (P // 2,) * 3, 64)
wt = torch.randn(64, 10, 3, 3, 3, device=DEV) * 0.05
wtf, wtd = ops.pack_weight(gt, _lib.W_CONVTR_FPROP, wt, DT), ops.pack_weight(gt, _lib.W_CONVTR_DGRAD, wt, DT)
dxt = ops.alloc_like(xt)
check("convT 64->10 fprop+stats", lambda: (ops.conv_fprop_stats(gt, xt, wtf, b, y), y))
check("convT 64->10 dgrad", lambda: ops.conv_dgrad(gt, dy, wtd, dxt))
check("convT 64->10 wgrad", lambda: ops.conv_wgrad(gt, xt, dy, want_bias=False))
g32 = ops.ConvGeom(3, 32, 32, 3, 1, False)
x32, dy32 = act(N, (P // 2,) * 3, 32), act(N, (P // 2,) * 3, 32)
w32 = torch.randn(32, 32, 3, 3, 3, device=DEV) * 0.05
check("conv 32->32 wgrad @80^3", lambda: ops.conv_wgrad(g32, x32, dy32, want_bias=False))
m32, r32 = ops.instnorm_stats(x32)
dx32 = ops.alloc_like(x32)
check("IN bwd 32ch @80^3", lambda: (ops.instnorm_prelu_bwd(x32, m32, r32, alpha, dy32, dx32), dx32))
