"""Two launches each of the kernels the second round-2 session changed, for an `ncu --set full` capture:
head conv 10->10 @128^3 x2 (fprop, dgrad: the two-CTAs-per-SM build; fprop with the fused residual: three CTAs) and the
cluster InstanceNorm+PReLU backward at 64 ch @16^3 x2 and 32 ch @32^3 x2.

    ncu --set full --clock-control none --import-source on -k regex:'tc_slide_conv|instnorm_prelu_bwd_cluster' \
        -o gpurun_out/r2b_kernels python scripts/profile_r2b.py
    python scripts/ncu_summary.py gpurun_out/r2b_kernels.ncu-rep profiles/r2b_ncu_full_kernels.csv
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ct_image_segmentation_b200 import _lib, ops

dev, dt = torch.device("cuda", 0), torch.bfloat16


def act(n, sp, c):
    t = ops.alloc_activation(n, sp, c, dt, dev)
    t.copy_(torch.randn(t.shape, device=dev))
    return t


g = ops.ConvGeom(3, 10, 10, 3, 1, False)
sp = (128, 128, 128)
x, dy, res = act(2, sp, 10), act(2, sp, 10), act(2, sp, 10)
y, dx = ops.alloc_like(x), ops.alloc_like(x)
w = torch.randn(10, 10, 3, 3, 3, device=dev) * 0.1
wf, wd = ops.pack_weight(g, _lib.W_CONV_FPROP, w, dt), ops.pack_weight(g, _lib.W_CONV_DGRAD, w, dt)
b = torch.zeros(10, device=dev)
for _ in range(2):
    ops.conv_fprop(g, x, wf, b, y)
for _ in range(2):
    ops.conv_dgrad(g, dy, wd, dx)
for _ in range(2):
    ops.conv_fprop(g, x, wf, b, y, res)
alpha = torch.full((1,), 0.25, device=dev)
for c, s in ((64, 16), (32, 32)):
    xc, dc = act(2, (s, s, s), c), act(2, (s, s, s), c)
    gx = ops.alloc_like(xc)
    mean, rstd = ops.instnorm_stats(xc)
    for _ in range(2):
        ops.instnorm_prelu_bwd(xc, mean, rstd, alpha, dc, gx)
torch.cuda.synchronize()
print("ok")
