"""Per-layer kernel timings at the bench shapes (CUDA events, no profiler):
    python scripts/kbench.py [case ...]      cases: head, l0, l1, convtr, convtr64, down, first, deep, norm, dice
Prints us per call for fprop / dgrad / wgrad of each layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ct_image_segmentation_b200 import _lib, ops
from ct_image_segmentation_b200.ops import ConvGeom

DEV = torch.device("cuda", 0)
DT = torch.bfloat16
lib = _lib.load()


def timeit(fn, reps=10, warm=3):
    """us per call, replayed from a CUDA graph (no host launch overhead in the number)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def conv_case(name, cin, cout, k, s, tr, n, sp_in, x_wide=None, stats=False):
    g = ConvGeom(3, cin, cout, k, s, tr)
    sp_out = g.out_spatial(*sp_in)
    if x_wide:  # x is the leading channel slice of a wider (concat) buffer
        xb = ops.alloc_activation(n, sp_in, x_wide, DT, DEV)
        xb.copy_(torch.randn(xb.shape, device=DEV))
        x = xb[..., :cin]
    else:
        x = ops.alloc_activation(n, sp_in, cin, DT, DEV)
        x.copy_(torch.randn(x.shape, device=DEV))
    y = ops.alloc_activation(n, sp_out, cout, DT, DEV)
    dy = ops.alloc_activation(n, sp_out, cout, DT, DEV)
    dy.copy_(torch.randn(dy.shape, device=DEV))
    dx = ops.alloc_like(x)
    ks = (k,) * 3
    w = torch.randn((cin, cout, *ks) if tr else (cout, cin, *ks), device=DEV) * 0.05
    kf = _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP
    kd = _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD
    wf, wd = ops.pack_weight(g, kf, w, DT), ops.pack_weight(g, kd, w, DT)
    b = torch.zeros(cout, device=DEV)
    out = []
    if stats:
        t = timeit(lambda: ops.conv_fprop_stats(g, x, wf, b, y))
    else:
        t = timeit(lambda: ops.conv_fprop(g, x, wf, b, y))
    out.append(("fprop", t, lib.b200seg_last_launch().decode()))
    if cin >= 8:
        t = timeit(lambda: ops.conv_dgrad(g, dy, wd, dx))
        out.append(("dgrad", t, lib.b200seg_last_launch().decode()))
    t = timeit(lambda: ops.conv_wgrad(g, x, dy, want_bias=False))
    out.append(("wgrad", t, lib.b200seg_last_launch().decode()))
    vox_in = n * sp_in[0] * sp_in[1] * sp_in[2]
    vox_out = n * sp_out[0] * sp_out[1] * sp_out[2]
    mb = (vox_in * max(cin, 8) + vox_out * ((cout + 15) // 16 * 16)) * 2 / 1e6
    if os.environ.get("KB_COMPACT"):
        print(f"{name}: " + " ".join(f"{a}={t:.1f}" for a, t, _ in out), end=" | ", flush=True)
        return
    print(f"{name:10s} {cin:3d}->{cout:3d} k{k} s{s} {'T' if tr else ' '} in {sp_in}: " +
          "  ".join(f"{a} {t:7.1f} us [{kn}]" for a, t, kn in out) + f"   (x+y = {mb:.0f} MB)")


CASES = {
    "head": lambda: conv_case("head", 10, 10, 3, 1, False, 2, (128, 128, 128)),
    "headL2": lambda: conv_case("headL2", 10, 10, 3, 1, False, 2, (32, 128, 128)),
    "col": lambda: conv_case("col", 27, 16, 1, 1, False, 2, (64, 64, 64), stats=True),
    "l0": lambda: conv_case("l0", 16, 16, 3, 1, False, 2, (64, 64, 64), stats=True),
    "l1": lambda: conv_case("l1", 32, 32, 3, 1, False, 2, (32, 32, 32), stats=True),
    "convtr": lambda: conv_case("convtr", 32, 10, 3, 2, True, 2, (64, 64, 64), stats=True),
    "convtr64": lambda: conv_case("convtr64", 64, 16, 3, 2, True, 2, (32, 32, 32), stats=True),
    "down": lambda: conv_case("down", 16, 32, 3, 2, False, 2, (64, 64, 64), stats=True),
    "down2": lambda: conv_case("down2", 32, 64, 3, 2, False, 2, (32, 32, 32), stats=True),
    "first": lambda: conv_case("first", 1, 16, 3, 2, False, 2, (128, 128, 128), stats=True),
    "deep": lambda: (conv_case("l2", 64, 64, 3, 1, False, 2, (16, 16, 16), stats=True),
                     conv_case("l3", 128, 128, 3, 1, False, 2, (8, 8, 8), stats=True),
                     conv_case("bot0", 128, 256, 3, 1, False, 2, (8, 8, 8), stats=True),
                     conv_case("bot1", 256, 256, 3, 1, False, 2, (8, 8, 8), stats=True),
                     conv_case("down3", 64, 128, 3, 2, False, 2, (16, 16, 16), stats=True),
                     conv_case("up3", 384, 64, 3, 2, True, 2, (8, 8, 8), stats=True),
                     conv_case("up2", 128, 32, 3, 2, True, 2, (16, 16, 16), stats=True)),
}

if __name__ == "__main__":
    torch.cuda.set_device(0)
    names = sys.argv[1:] or list(CASES)
    for nme in names:
        CASES[nme]()
