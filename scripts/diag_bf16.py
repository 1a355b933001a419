"""Diagnostic (GPU): where do bf16 GPU gradients diverge from the bf16-rounding CPU evaluation?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ct_image_segmentation_b200 as B
import ct_image_segmentation_b200.unet as U
from ct_image_segmentation_b200 import ops as real_ops
from oracle import monai_ref as O
from tests import _torch_ops

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

torch.manual_seed(12342)
ch, st, res, shape = [16, 32, 64, 128, 256], [2, 2, 2, 2], 2, (2, 1, 32, 48, 32)
ref = O.UNet(3, 1, 10, ch, st, num_res_units=res)
torch.manual_seed(1)
x = torch.randn(*shape)
lab = torch.randint(0, 10, (shape[0], *shape[2:]))
dt = torch.bfloat16 if len(sys.argv) < 2 else torch.float32
emu = B.UNet(3, 1, 10, ch, st, num_res_units=res, dtype=dt); emu.load_state_dict(ref.state_dict())
net = B.UNet(3, 1, 10, ch, st, num_res_units=res, dtype=dt); net.load_state_dict(ref.state_dict()); net = net.cuda()

U.ops = _torch_ops
saved_e = {}
out_e = emu._run_forward(_torch_ops.to_channels_last(x, dt), saved_e, keep_all=False)
fw_e = {m: {k: (v.clone() if torch.is_tensor(v) else v) for k, v in s.items()} for m, s in saved_e.items()}
le = _torch_ops.from_channels_last(out_e, 3).float().requires_grad_(True)
O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(le, lab.unsqueeze(1)).backward()
ge = le.grad.to(dt)
taps_e = {}
grads_e, _ = emu._run_backward(saved_e, _torch_ops.to_channels_last(ge, dt), False, taps_e)

U.ops = real_ops
saved_g = {}
out_g = net._run_forward(real_ops.to_channels_last(x.cuda(), dt), saved_g)
fw_g = dict(saved_g)
print("logits", rel(out_g, out_e))
lg = real_ops.from_channels_last(out_g, 3).detach().requires_grad_(True)
B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(lg, lab.cuda().unsqueeze(1)).backward()
print("dlogits gpu-kernel vs oracle-autograd(rounded)", rel(lg.grad, ge.float()))
# feed the SAME dlogits (emulation's) to the GPU backward to isolate the backward kernels
taps_g = {}
g_in = real_ops.to_channels_last(ge.cuda(), dt)
grads_g, _ = net._run_backward(saved_g, g_in, False, taps_g)
names_e = {m: n for n, m in emu.named_modules()}
names_g = {m: n for n, m in net.named_modules()}
mods_g = {n: m for m, n in names_g.items()}
print("--- forward saved tensors (c, mean, rstd) per conv")
for m_e, s in fw_e.items():
    n = names_e[m_e]
    if not isinstance(m_e, U.Convolution): continue
    sg = fw_g[mods_g[n]]
    if s.get("c") is not None:
        print(f"{n:70s} c {rel(sg['c'], s['c']):.2e} ")
print("--- backward taps d/d conv-out, in backward order")
for m_e, t in taps_e.items():
    n = names_e[m_e]
    tg = taps_g[mods_g[n]]['g_c']; t = t['g_c']
    pe = dict(emu.named_parameters()); pg = dict(net.named_parameters())
    we = rel(grads_g[pg[n + '.conv.weight']], grads_e[pe[n + '.conv.weight']])
    print(f"{n:70s} g_c {rel(tg, t):.2e}  |g_c| {t.float().abs().mean().item():.2e} mean/absmean {abs(t.float().mean().item())/t.float().abs().mean().item():.2e} gw {we:.2e}")
