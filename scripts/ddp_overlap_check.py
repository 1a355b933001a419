"""Multi-GPU check of GraphedTrainStep's gradient exchange (run under torchrun, one rank per GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      scripts/ddp_overlap_check.py

Every rank trains on its own batch.  Checked on every rank:
  * the all-reduced bucket of the OVERLAPPED exchange (deep range reduced beside the backward pass, inside the
    CUDA graph) equals the bucket of the plain exchange (one all-reduce after the backward pass) -- bit for bit
    with 2 ranks, 1e-6 otherwise (NCCL's reduction order may differ between message sizes);
  * it equals the mean over ranks of the local gradients (all-gathered), 1e-6;
  * replays keep giving the same bucket.
Prints one line per rank and exits non-zero on a mismatch.
"""
import gc
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ct_image_segmentation_b200 as B  # noqa: E402
from ct_image_segmentation_b200.parallel import init_distributed  # noqa: E402


def main():
    rank, world, local = init_distributed("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    torch.manual_seed(12342)
    net = B.UNet(3, 1, 10, [16, 32, 64, 128, 256], [2, 2, 2, 2], num_res_units=2, dtype=torch.bfloat16).to(dev)
    fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    g = torch.Generator().manual_seed(100 + rank)
    x = torch.randn(2, 1, 48, 48, 48, generator=g).to(dev)
    lab = torch.randint(0, 10, (2, 48, 48, 48), generator=g).to(dev)

    # local gradients (no exchange): plain autograd
    fx(net(x), lab.unsqueeze(1)).backward()
    local_flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    for p in net.parameters():
        p.grad = None
    gathered = [torch.empty_like(local_flat) for _ in range(world)]
    dist.all_gather(gathered, local_flat)
    mean_flat = torch.stack(gathered).double().mean(0).float()

    results = {}
    for overlap in (False, True):
        for use_graph in (False, True):
            print(f"rank {rank}: overlap={overlap} graph={use_graph} ...", file=sys.stderr, flush=True)
            train = B.GraphedTrainStep(net, fx, None, x, lab, use_graph=use_graph, overlap_allreduce=overlap)
            assert (train._deep is not None) == overlap
            outs = []
            for _ in range(3):
                train(None, None)
                torch.cuda.synchronize()
                outs.append(train.bucket.flat.clone())
            assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "replays differ"
            results[(overlap, use_graph)] = outs[0]
            net.enable_wgrad_stream(False)
            # a CUDA graph holding NCCL kernels must be gone before the next communicator use / teardown
            del train
            gc.collect()
            torch.cuda.synchronize()
            dist.barrier()
    base = results[(False, False)]
    ok = True
    for key, val in results.items():
        err = ((val - mean_flat).norm() / mean_flat.norm()).item()
        same = torch.equal(val, base)
        close = ((val - base).norm() / base.norm()).item()
        print(f"rank {rank}: overlap={key[0]} graph={key[1]} vs mean-of-local {err:.2e}, vs plain exchange "
              f"{'bit-equal' if same else f'{close:.2e}'}", flush=True)
        ok = ok and err < 1e-6 and (same if world == 2 else close < 1e-6)
    t = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(t)
    bad = int(t.item())
    if rank == 0 and not bad:
        print("ddp overlap check OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if bad:
        sys.exit(1)


if __name__ == "__main__":
    main()
