"""Same-box GPU bar: the reference's path -- the oracle's MONAI-0.3 UNet + DiceLoss on
stock torch / cuDNN kernels -- timed ON THE B200 for BASELINE.json configs[2] (128^3, batch 2), as the
reference itself would run it (fp32 NCDHW with TF32 convolutions) and with the usual eager tuning (bf16
autocast + channels_last_3d).  Not a test and not part of the product: run as

    python -m tests.gpu_bar            # prints one JSON line

It lives under tests/ because only tests/ (and bench.py's CPU baseline legs) may execute oracle/.
"""
import json
import sys

import torch

from oracle import monai_ref as O


def timed_steps(net, images, labels, autocast, steps=5, warmup=2):
    loss_fx = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    times = []
    for i in range(warmup + steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            out = net(images)
        loss = loss_fx(out.float(), labels.unsqueeze(1))
        loss.backward()
        opt.step()
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    return min(times), sum(times) / len(times), float(loss.detach())


def main():
    assert torch.cuda.is_available()
    patch, batch = 128, 2
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(12342)
    images = torch.randn(batch, 1, patch, patch, patch, device="cuda")
    labels = torch.randint(0, 10, (batch, patch, patch, patch), device="cuda")
    vox = batch * patch ** 3
    out = {"workload": f"oracle MONAI-0.3 UNet 16-256 + DiceLoss + Adam, {patch}^3 x {batch}, torch {torch.__version__} eager on "
                       f"{torch.cuda.get_device_name(0)} (cudnn.benchmark)"}
    for tag, autocast, cl in (("fp32_tf32_ncdhw", False, False), ("bf16_autocast_channels_last_3d", True, True)):
        torch.manual_seed(12342)
        net = O.UNet(3, 1, 10, [16, 32, 64, 128, 256], [2, 2, 2, 2], num_res_units=2).cuda()
        x = images
        if cl:
            net = net.to(memory_format=torch.channels_last_3d)
            x = images.contiguous(memory_format=torch.channels_last_3d)
        try:
            best, mean, loss = timed_steps(net, x, labels, autocast)
            out[tag] = {"ms_per_step_best": best, "ms_per_step_mean": mean, "voxels_per_s": vox / (best * 1e-3), "loss": loss}
        except Exception as e:  # noqa: BLE001 -- report and go on with the next variant
            out[tag] = {"error": repr(e)[:300]}
        del net
        torch.cuda.empty_cache()
    print(json.dumps(out))
    return 0


if __name__ == "__main__":
    sys.exit(main())
