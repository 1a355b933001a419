"""world_size-2 gloo test of the data-parallel plumbing (CPU, no GPU kernels involved)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ct_image_segmentation_b200.parallel import GradientBucket, shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.PReLU(), torch.nn.Linear(5, 3))
    x = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0
    xs = x[rank * 2:(rank + 1) * 2]             # each rank: its own half of the global batch
    lin(xs).square().mean().backward()
    bucket = GradientBucket(lin.parameters())
    bucket.allreduce_mean()
    if rank == 0:
        out.put([p.grad.clone() for p in lin.parameters()])
    dist.barrier()
    dist.destroy_process_group()


def test_gradient_bucket_matches_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.PReLU(), torch.nn.Linear(5, 3))
    x = torch.arange(24, dtype=torch.float32).reshape(4, 6) / 10.0
    lin(x).square().mean().backward()           # mean over the global batch == mean of rank means
    for g, p in zip(got, lin.parameters()):
        torch.testing.assert_close(g, p.grad, rtol=1e-5, atol=1e-6)


def test_shard_indices_partition():
    items = 50
    parts = [shard_indices(items, r, 8) for r in range(8)]
    assert sorted(sum(parts, [])) == list(range(items))
    assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
