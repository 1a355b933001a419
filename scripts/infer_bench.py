"""cfg4 (BASELINE.json configs[3]): sliding-window inference on a synthetic 512x512x160 CT volume, ROI 128^3,
overlap 0.25 (50 windows), bf16, to a uint8 label map.  One GPU here (`--world N` under torchrun shards the
windows round-robin and all-reduces the accumulators).  Prints one JSON line: volume voxels/s and latency,
eager predictor vs. CUDA-graph predictor."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_image_segmentation_b200 as B  # noqa: E402
from ct_image_segmentation_b200.inference import GraphedPredictor, sliding_window_inference, window_list  # noqa: E402
from ct_image_segmentation_b200.parallel import init_distributed  # noqa: E402


def main():
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(12342)
    net = B.UNet(3, 1, 10, [16, 32, 64, 128, 256], [2, 2, 2, 2], num_res_units=2, dtype=torch.bfloat16).to(dev).eval()
    vol = torch.randn(1, 1, 160, 512, 512, device=dev)
    roi, swb = (128, 128, 128), 2
    nwin = len(window_list(vol.shape[2:], roi, 0.25))
    graphed = GraphedPredictor(net, torch.zeros(swb, 1, *roi, device=dev))

    def timed(pred, reps=3):
        best, lab = 1e30, None
        for _ in range(reps + 1):  # first pass = warm-up
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lab = sliding_window_inference(vol, roi, swb, pred, overlap=0.25)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best, lab

    ms_e, lab_e = timed(net)
    ms_g, lab_g = timed(graphed)
    vox = vol.numel()
    if rank == 0:
        print(json.dumps({
            "workload": "sliding-window inference, 512x512x160 volume, roi 128^3, overlap 0.25, sw_batch 2, bf16",
            "windows": nwin, "n_gpus": world,
            "eager_ms": ms_e, "eager_voxels_per_s": vox / (ms_e * 1e-3),
            "graph_ms": ms_g, "graph_voxels_per_s": vox / (ms_g * 1e-3),
            "labels_equal": bool(torch.equal(lab_e, lab_g)), "label_dtype": str(lab_g.dtype),
            "label_shape": list(lab_g.shape)}))


if __name__ == "__main__":
    main()
