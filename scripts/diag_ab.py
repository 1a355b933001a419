import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ct_image_segmentation_b200 as B
import ct_image_segmentation_b200.unet as U
from ct_image_segmentation_b200 import ops, _lib
from oracle import monai_ref as O
from tests.test_gpu_unet import make_pair, sparse_labels, rel

dt = torch.float32
ref, netA = make_pair(3, 1, [16, 32, 64, 128, 256], [2, 2, 2, 2], 2, dt)
_, netB = make_pair(3, 1, [16, 32, 64, 128, 256], [2, 2, 2, 2], 2, dt)
torch.manual_seed(1)
shape = (2, 1, 32, 48, 32)
x = torch.randn(*shape); lab = sparse_labels(shape[0], shape[2:])
xc = ops.to_channels_last(x.cuda(), dt)

def run(net, force):
    orig_f = ops.conv_fprop
    if force:
        ops.conv_fprop = lambda g, x, wp, b, y, residual=None, flags=0: orig_f(g, x, wp, b, y, residual, flags | _lib.CONV_FORCE_GENERIC)
    saved = {}
    out = net._run_forward(xc, saved)
    ops.conv_fprop = orig_f
    fw = dict(saved)
    lg = ops.from_channels_last(out, 3).detach().requires_grad_(True)
    B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(lg, lab.cuda().unsqueeze(1)).backward()
    taps = {}
    grads, _ = net._run_backward(saved, ops.to_channels_last(lg.grad, dt), False, taps)
    return fw, taps, grads, lg.grad

fwA, tA, gA, dlA = run(netA, False)
fwB, tB, gB, dlB = run(netB, True)
nA = {m: n for n, m in netA.named_modules()}; mB = {n: m for n, m in netB.named_modules()}
print("dlogits A vs B", rel(dlA, dlB))
for m, s in fwA.items():
    if not isinstance(m, U.Convolution): continue
    sb = fwB[mB[nA[m]]]
    msg = f"{nA[m]:66s}"
    for k in ("x", "c", "mean", "rstd"):
        if s.get(k) is not None: msg += f" {k} {rel(s[k], sb[k]):.1e}"
    print(msg)
print("--- backward")
for m, t in tA.items():
    tb = tB[mB[nA[m]]]
    print(f"{nA[m]:66s} g_out {rel(t['g_out'], tb['g_out']):.1e} g_c {rel(t['g_c'], tb['g_c']):.1e}")
