"""Which layer's output differs between two identical forward passes?  python scripts/determinism_fwd.py [batch] [patch] [side]"""
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
x = torch.randn(batch, 1, patch, patch, patch, generator=torch.Generator().manual_seed(1)).to(dev)
runs = []
for _ in range(3):
    taps = net.forward_debug(x)
    torch.cuda.synchronize()
    runs.append({k: v.clone() for k, v in taps.items()})
    del taps
print(f"batch {batch} patch {patch} side {side}")
for k in runs[0]:
    eq = [bool(torch.equal(runs[0][k], r[k])) for r in runs[1:]]
    if not all(eq):
        d = max((runs[0][k] - r[k]).abs().max().item() for r in runs[1:])
        n = max(int((runs[0][k] != r[k]).sum()) for r in runs[1:])
        print(f"  DIFF {k or 'logits'}: max abs {d:.3e}, {n} of {runs[0][k].numel()} elements")
print("done")
