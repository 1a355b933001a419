"""Eager forward + Dice + backward twice with identical inputs: which parameter gradients differ?
python scripts/determinism_bwd.py [batch] [patch] [side]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ct_image_segmentation_b200 as B  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4
patch = int(sys.argv[2]) if len(sys.argv) > 2 else 160
side = int(sys.argv[3]) if len(sys.argv) > 3 else 0
filters = [32, 64, 128, 256, 512]
dev = torch.device("cuda", 0)
torch.manual_seed(12342)
net = B.UNet(3, 1, 10, filters, [2, 2, 2, 2], num_res_units=2, dtype=torch.bfloat16).to(dev)
net.enable_wgrad_stream(bool(side))
g = torch.Generator().manual_seed(1)
x = torch.randn(batch, 1, patch, patch, patch, generator=g).to(dev)
lab = torch.randint(0, 10, (batch, 1, patch, patch, patch), generator=g, dtype=torch.uint8).to(dev)
fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
runs = []
for _ in range(3):
    for p in net.parameters():
        p.grad = None
    y = net(x)
    loss = fx(y, lab)
    loss.backward()
    torch.cuda.synchronize()
    runs.append(([p.grad.clone() for p in net.parameters()], y.detach().clone(), loss.item()))
print(f"batch {batch} patch {patch} side {side}; losses {[r[2] for r in runs]}; logits equal "
      f"{[bool(torch.equal(runs[0][1], r[1])) for r in runs[1:]]}")
names = [n for n, _ in net.named_parameters()]
nd = 0
for i, n in enumerate(names):
    eq = [bool(torch.equal(runs[0][0][i], r[0][i])) for r in runs[1:]]
    if not all(eq):
        nd += 1
        if nd <= 12:
            d = max((runs[0][0][i] - r[0][i]).abs().max().item() for r in runs[1:])
            print(f"  DIFF {n}: max abs {d:.3e} (max |g| {runs[0][0][i].abs().max().item():.3e})")
print(f"{nd} of {len(names)} gradients differ")
