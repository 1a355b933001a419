"""Times the head layer (conv 10->10, 3x3x3, stride 1, 128^3 x batch) fprop / dgrad / wgrad alone,
CUDA events on the launching stream.  Target of the `ncu --set full` captures in profiles/."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ct_image_segmentation_b200 import ops, _lib

cin = cout = int(sys.argv[1]) if len(sys.argv) > 1 else 10
p = int(sys.argv[2]) if len(sys.argv) > 2 else 128
n = 2
dt = torch.bfloat16
g = ops.ConvGeom(3, cin, cout, 3, 1, False)
x = ops.alloc_activation(n, (p, p, p), cin, dt, "cuda"); x.copy_(torch.randn(x.shape, device="cuda"))
y = ops.alloc_activation(n, (p, p, p), cout, dt, "cuda")
dy = ops.alloc_activation(n, (p, p, p), cout, dt, "cuda"); dy.copy_(torch.randn(dy.shape, device="cuda"))
dx = ops.alloc_like(x)
w = torch.randn(cout, cin, 3, 3, 3, device="cuda") * 0.1
wf = ops.pack_weight(g, _lib.W_CONV_FPROP, w, dt); wd = ops.pack_weight(g, _lib.W_CONV_DGRAD, w, dt)
b = torch.zeros(cout, device="cuda")
def timeit(f, reps=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
vox = n * p ** 3
flop = 2.0 * 27 * cin * cout * vox
for name, f in (("fprop", lambda: ops.conv_fprop(g, x, wf, b, y)),
                ("dgrad", lambda: ops.conv_dgrad(g, dy, wd, dx)),
                ("wgrad", lambda: ops.conv_wgrad(g, x, dy, want_bias=False))):
    us = timeit(f)
    print(f"{name}: {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s  ({vox * 2 * 16 * 2 / us / 1e3:6.0f} GB/s of 16-ch rows in+out)")
