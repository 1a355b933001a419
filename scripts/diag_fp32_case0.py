import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ct_image_segmentation_b200 as B
from ct_image_segmentation_b200 import ops, _lib
from oracle import monai_ref as O
from tests.test_gpu_unet import make_pair, sparse_labels, rel, dead_bias

def run(tag):
    ref, net = make_pair(3, 1, [16, 32, 64, 128, 256], [2, 2, 2, 2], 2, torch.float32)
    torch.manual_seed(1)
    shape = (2, 1, 32, 48, 32)
    x = torch.randn(*shape); lab = sparse_labels(shape[0], shape[2:])
    y_ref = ref(x)
    O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y_ref, lab.unsqueeze(1)).backward()
    y = net(x.cuda())
    B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(y, lab.cuda().unsqueeze(1)).backward()
    rp = dict(ref.named_parameters())
    errs = {n: rel(p.grad, rp[n].grad) for n, p in net.named_parameters() if not dead_bias(n, rp)}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(tag, "logits", rel(y, y_ref.detach()), "worst", [(n[-40:], f"{e:.2e}") for n, e in worst])

run("current")
orig_w = ops.conv_wgrad
ops.conv_wgrad = lambda g, x, dy, want_bias=True, flags=0: orig_w(g, x, dy, True, flags)
run("all-bias-colsum")
ops.conv_wgrad = lambda g, x, dy, want_bias=True, flags=0: orig_w(g, x, dy, want_bias, flags | _lib.CONV_FORCE_GENERIC)
run("generic-wgrad")
ops.conv_wgrad = orig_w
orig_f = ops.conv_fprop
ops.conv_fprop = lambda g, x, wp, b, y, residual=None, flags=0: orig_f(g, x, wp, b, y, residual, flags | _lib.CONV_FORCE_GENERIC)
run("generic-fprop")
