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


# ---- the U-Net's overlapped gradient exchange (engine.GraphedTrainStep), two gloo ranks ------------------
def _unet_worker(rank, world, port, out):
    import types

    import ct_image_segmentation_b200.unet as U
    from ct_image_segmentation_b200.engine import GraphedTrainStep
    from oracle import monai_ref as O

    from . import _torch_ops

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    U.ops = _torch_ops                       # kernels replaced by their torch emulation (CPU)
    torch.manual_seed(12342)
    ch, st = [4, 8, 8, 16, 16], [2, 2, 2, 2]
    ref = O.UNet(3, 1, 10, ch, st, num_res_units=2)
    net = U.UNet(3, 1, 10, ch, st, num_res_units=2, dtype=torch.float32)
    net.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(2, 1, 32, 32, 32, generator=g)
    lab_all = torch.randint(0, 10, (2, 32, 32, 32), generator=g)
    x, lab = x_all[rank:rank + 1], lab_all[rank:rank + 1]        # each rank: its own sample

    # the loss gradient w.r.t. the logits comes from the oracle's Dice loss on this rank's sample
    saved = {}
    logits_cl = net._run_forward(_torch_ops.to_channels_last(x, torch.float32), saved)
    logits = _torch_ops.from_channels_last(logits_cl, 3).detach().requires_grad_(True)
    O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(logits, lab.unsqueeze(1)).backward()
    g_out = _torch_ops.to_channels_last(logits.grad, torch.float32)

    bucket = GradientBucket(net.parameters())
    sink = dict(zip(bucket.params, bucket.views))
    depth, lo, hi = GraphedTrainStep._deep_range(types.SimpleNamespace(bucket=bucket), net)
    assert depth == 2 and 0 < lo < hi == bucket.flat.numel()
    events = []

    def level_done(d):                       # the engine's _on_level_done, with gloo instead of NCCL + streams
        if d != depth:
            return
        # gradients that are not complete yet must not be touched by the early exchange
        events.append(bucket.flat[:lo].clone())
        dist.all_reduce(bucket.flat[lo:hi])
        bucket.flat[lo:hi].mul_(1.0 / world)

    net.bind_grad_sink(sink, level_done)
    net._run_backward(saved, g_out, False)
    net.bind_grad_sink(None)
    assert len(events) == 1
    dist.all_reduce(bucket.flat[:lo])        # the engine's _exchange_rest
    bucket.flat[:lo].mul_(1.0 / world)
    if rank == 0:
        out.put((bucket.flat.clone(), [p.numel() for p in bucket.params], lo))
    dist.barrier()
    dist.destroy_process_group()


def test_unet_overlapped_exchange_matches_global_batch():
    """Two gloo ranks run the U-Net's reverse plan (kernels emulated) with the gradient sink and the level
    hook: early all-reduce of the range [lo, end) when level 2 is done, the first range afterwards.  The
    bucket then holds the gradient of the GLOBAL batch (oracle autograd on both samples, Dice `mean`)."""
    from oracle import monai_ref as O
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_unet_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    flat, sizes, lo = out.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    torch.manual_seed(12342)
    ref = O.UNet(3, 1, 10, [4, 8, 8, 16, 16], [2, 2, 2, 2], num_res_units=2)
    g = torch.Generator().manual_seed(7)
    x_all = torch.randn(2, 1, 32, 32, 32, generator=g)
    lab_all = torch.randint(0, 10, (2, 32, 32, 32), generator=g)
    O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)(ref(x_all), lab_all.unsqueeze(1)).backward()
    assert 0 < lo < flat.numel()
    names = [n for n, _ in ref.named_parameters()]
    for name, p, got in zip(names, ref.parameters(), flat.split(sizes)):
        want = p.grad.reshape(-1)
        if name.endswith("conv.bias") and name[:-len("conv.bias")] + "act.weight" in names:
            assert float(got.abs().max()) == 0.0, name       # dead bias: exactly zero, never exchanged noise
            continue
        scale = max(float(want.abs().max()), 1e-6)
        assert float((got - want).abs().max()) <= 5e-3 * scale + 1e-6, name
