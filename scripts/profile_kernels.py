"""One launch (after a warm-up launch) of each kernel the step's time is made of, at the cfg3 shapes (128^3 x 2,
bf16): the target of the `ncu --set full` captures kept under profiles/ (tensor-pipe utilisation
for the convs, achieved HBM GB/s and DRAM traffic for the norm / loss kernels).

    python scripts/profile_kernels.py && ncu --set full --clock-control none --import-source on \
        -k regex:'tc_slide|instnorm_prelu_bwd|softmax_dice|tc_convtr_fprop|tc_conv_splitk|tc_wgrad_kernel' \
        -o gpurun_out/r2_kernels python scripts/profile_kernels.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ct_image_segmentation_b200 import _lib, ops  # noqa: E402

DEV, DT = torch.device("cuda", 0), torch.bfloat16
N, P = 2, 128


def act(n, sp, c):
    t = ops.alloc_activation(n, sp, c, DT, DEV)
    t.copy_(torch.randn(t.shape, device=DEV))
    return t


def conv(cin, cout, k, s, tr, n, sp_in, what):
    g = ops.ConvGeom(3, cin, cout, k, s, tr)
    sp_out = g.out_spatial(*sp_in)
    x, dy = act(n, sp_in, cin), act(n, sp_out, cout)
    y, dx = ops.alloc_activation(n, sp_out, cout, DT, DEV), ops.alloc_like(x)
    w = torch.randn((cin, cout, k, k, k) if tr else (cout, cin, k, k, k), device=DEV) * 0.05
    wf = ops.pack_weight(g, _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP, w, DT)
    wd = ops.pack_weight(g, _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD, w, DT)
    b = torch.zeros(cout, device=DEV)
    runs = {"fprop": lambda: ops.conv_fprop(g, x, wf, b, y), "dgrad": lambda: ops.conv_dgrad(g, dy, wd, dx),
            "wgrad": lambda: ops.conv_wgrad(g, x, dy, want_bias=False),
            "fprop_stats": lambda: ops.conv_fprop_stats(g, x, wf, b, y)}
    for op in what:
        for _ in range(2):
            runs[op]()
        torch.cuda.synchronize()
        print(f"{cin}->{cout} k{k} s{s}{'T' if tr else ''} {op}: {_lib.load().b200seg_last_launch().decode()}")


def main():
    torch.cuda.set_device(0)
    conv(10, 10, 3, 1, False, N, (P, P, P), ("fprop", "wgrad"))
    conv(32, 10, 3, 2, True, N, (P // 2,) * 3, ("fprop_stats",))
    conv(16, 16, 3, 1, False, N, (P // 2,) * 3, ("fprop_stats", "wgrad"))
    conv(256, 256, 3, 1, False, N, (P // 16,) * 3, ("fprop_stats", "wgrad"))
    conv(64, 64, 3, 1, False, N, (P // 8,) * 3, ("fprop_stats", "wgrad"))
    # InstanceNorm + PReLU on the full-resolution 10-class tensor
    x, dy = act(N, (P, P, P), 10), act(N, (P, P, P), 10)
    y, dx = ops.alloc_like(x), ops.alloc_like(x)
    alpha = torch.full((1,), 0.25, device=DEV)
    mean, rstd = ops.instnorm_stats(x)
    for _ in range(2):
        ops.instnorm_prelu_fwd(x, mean, rstd, alpha, y)
        ops.instnorm_prelu_bwd(x, mean, rstd, alpha, dy, dx)
    # softmax + Dice (+ metric counts), forward and backward
    lab = torch.randint(0, 10, (N, P, P, P), device=DEV, dtype=torch.uint8)
    gi = torch.rand(N, 10, device=DEV)
    for _ in range(2):
        ops.softmax_dice_metric_sums(x, lab)
        ops.softmax_dice_bwd(x, lab, gi, gi)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
