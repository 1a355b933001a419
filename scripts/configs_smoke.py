"""Runs every BASELINE.json configuration once at full size on the GPU (bf16, fwd + Dice + bwd) and
prints ms/step and voxels/s: a robustness pass (ragged tiles, 512-channel layers, 2-D path)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ct_image_segmentation_b200 as B
from ct_image_segmentation_b200 import _lib

lib = _lib.load()
CFG = [
    ("cfg1 2D 64-1024 512^2 b4", 2, [64, 128, 256, 512, 1024], (4, 1, 512, 512)),
    ("cfg2 3D 16-256 96^3 b2", 3, [16, 32, 64, 128, 256], (2, 1, 96, 96, 96)),
    ("cfg3 3D 16-256 128^3 b1", 3, [16, 32, 64, 128, 256], (1, 1, 128, 128, 128)),
    ("cfg5 3D 32-512 160^3 b2", 3, [32, 64, 128, 256, 512], (2, 1, 160, 160, 160)),
]
for name, dims, ch, shape in CFG:
    torch.manual_seed(0)
    net = B.UNet(dims, 1, 10, ch, [2, 2, 2, 2], num_res_units=2).cuda()
    loss_fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    x = torch.randn(*shape, device="cuda")
    lab = torch.randint(0, 10, (shape[0], *shape[2:]), device="cuda", dtype=torch.uint8)
    def step():
        for p in net.parameters(): p.grad = None
        l = loss_fx(net(x), lab.unsqueeze(1)); l.backward(); return l
    t0 = lib.b200seg_tc_launch_count(); l0 = lib.b200seg_launch_count()
    loss = step(); torch.cuda.synchronize()
    tc, al = lib.b200seg_tc_launch_count() - t0, lib.b200seg_launch_count() - l0
    step(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    vox = shape[0] * int(torch.tensor(shape[2:]).prod())
    finite = all(torch.isfinite(p.grad).all().item() for p in net.parameters())
    print(f"{name}: loss {loss.item():.4f} grads finite {finite}  {ms:8.2f} ms/step (eager)  {vox / ms * 1e3:.3e} voxels/s  "
          f"launches {al} (tcgen05 {tc})  mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del net, x, lab
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
