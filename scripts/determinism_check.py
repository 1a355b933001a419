"""Run-to-run determinism of the training step on ONE GPU: graph replay vs graph replay, and graph replay vs the
eager issue of the same launches (GraphedTrainStep.check_exchange with world 1).   python scripts/determinism_check.py cfg5"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ct_image_segmentation_b200 as B  # noqa: E402

CFG = {"cfg3": ([16, 32, 64, 128, 256], 128, 2), "cfg5": ([32, 64, 128, 256, 512], 160, 4),
       "cfg5b1": ([32, 64, 128, 256, 512], 160, 1), "cfg5_64": ([32, 64, 128, 256, 512], 64, 2)}


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    filters, patch, batch = CFG[name]
    dev = torch.device("cuda", 0)
    torch.manual_seed(12342)
    net = B.UNet(3, 1, 10, filters, [2, 2, 2, 2], num_res_units=2, dtype=torch.bfloat16).to(dev)
    fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, with_metric=True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(batch, 1, patch, patch, patch, generator=g).to(dev)
    lab = torch.randint(0, 10, (batch, patch, patch, patch), generator=g, dtype=torch.uint8).to(dev)
    train = B.GraphedTrainStep(net, fx, None, x, lab, metric="fused")
    names = [n for n, _ in net.named_parameters()]
    train(None, None)
    torch.cuda.synchronize()
    a = train.bucket.flat.clone()
    train(None, None)
    torch.cuda.synchronize()
    b = train.bucket.flat.clone()
    e = train.local_gradients()
    torch.cuda.synchronize()
    print(name, "replay vs replay bitwise:", bool(torch.equal(a, b)), " replay vs eager bitwise:", bool(torch.equal(b, e)))
    for tag, u, v in (("replay/replay", a, b), ("replay/eager", b, e)):
        if torch.equal(u, v):
            continue
        for n, pu, pv in zip(names, u.split(train.bucket.sizes), v.split(train.bucket.sizes)):
            if not torch.equal(pu, pv):
                d = (pu - pv).abs().max().item()
                print(f"  {tag}: {n}: max abs diff {d:.3e} (max |g| {pv.abs().max().item():.3e})")


if __name__ == "__main__":
    main()
