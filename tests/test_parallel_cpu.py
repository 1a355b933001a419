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


# ---- sliding-window inference sharded over ranks (inference.py), gloo ---------------------------------------
class _TorchWindowOps:
    """torch emulation of b200seg_window_accumulate_weighted / b200seg_accum_argmax (host-logic test on CPU)."""

    def check(self, x):
        pass

    def accumulate(self, piece, imp, acc, cnt, d0, h0, w0):
        r, h, w, _ = piece.shape
        a, c = acc[d0:d0 + r, h0:h0 + h, w0:w0 + w], cnt[d0:d0 + r, h0:h0 + h, w0:w0 + w]
        if imp is None:
            a += piece.float()
            c += 1.0
        else:
            a += piece.float() * imp[..., None]
            c += imp

    def argmax(self, acc, cnt, want_mean):
        mean = acc / cnt[..., None]
        return torch.softmax(mean, -1).argmax(-1).to(torch.uint8), (mean if want_mean else None)


def _toy_predictor():
    torch.manual_seed(3)
    net = torch.nn.Sequential(torch.nn.Conv3d(1, 6, 3, padding=1), torch.nn.Tanh(), torch.nn.Conv3d(6, 10, 1))
    net.out_channels = 10
    return net.eval()


def _infer_worker(rank, world, port, out, mode):
    from ct_image_segmentation_b200.inference import sliding_window_inference
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    net = _toy_predictor()
    x = torch.randn(1, 1, 30, 28, 20, generator=torch.Generator().manual_seed(9))
    lab, logits = sliding_window_inference(x, (16, 16, 16), 2, net, overlap=0.25, mode=mode, return_logits=True,
                                           _dev_ops=_TorchWindowOps())
    out.put((rank, lab, logits))
    dist.barrier()
    dist.destroy_process_group()


def _run_infer(world, mode):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_infer_worker, args=(r, world, port, out, mode)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict((r, (lab, lg)) for r, lab, lg in (out.get(timeout=300) for _ in range(world)))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    return got


def test_sharded_sliding_window_two_and_three_ranks():
    """World 2 and 3 over gloo (uneven slabs: 30 rows / 3 ranks, 18 windows): every rank ends with the whole label
    map; label map and averaged logits equal the single-process run BIT FOR BIT (accumulation in global window
    order at the slab owners) and the oracle's sliding_window_inference to rounding."""
    from ct_image_segmentation_b200.inference import make_plan, sliding_window_inference
    from oracle import monai_ref as O
    net = _toy_predictor()
    x = torch.randn(1, 1, 30, 28, 20, generator=torch.Generator().manual_seed(9))
    for world, mode in ((2, "constant"), (3, "gaussian")):
        single, single_logits = sliding_window_inference(x, (16, 16, 16), 2, net, overlap=0.25, mode=mode,
                                                         return_logits=True, rank=0, world=1,
                                                         _dev_ops=_TorchWindowOps())
        with torch.no_grad():
            want = O.sliding_window_inference(x, (16, 16, 16), 2, net, overlap=0.25, mode=mode)
        torch.testing.assert_close(single_logits, want, rtol=1e-5, atol=1e-6)
        got = _run_infer(world, mode)
        for r in range(world):
            lab, logits = got[r]
            assert lab.dtype == torch.uint8 and tuple(lab.shape) == (1, 30, 28, 20)
            assert torch.equal(lab, single) and torch.equal(logits, single_logits), (world, r)
        plan = make_plan((30, 28, 20), (16, 16, 16), 0.25, world)
        assert plan.bounds[-1] == 30 and plan.runs[-1] == len(plan.wins) == 3 * 2 * 2
        assert any(p.src != p.dst for p in plan.pieces)  # rows really cross rank borders in this case


def test_shard_plan_cfg4_geometry():
    """cfg4 (BASELINE.json configs[3]): 512x512x160 volume, ROI 128^3, overlap 0.25 on 8 ranks."""
    from ct_image_segmentation_b200.inference import make_plan
    plan = make_plan((512, 512, 160), (128, 128, 128), 0.25, 8)
    assert len(plan.wins) == 50
    per_rank = [plan.runs[r + 1] - plan.runs[r] for r in range(8)]
    assert sum(per_rank) == 50 and max(per_rank) == 7 and min(per_rank) == 6
    assert [plan.bounds[r + 1] - plan.bounds[r] for r in range(8)] == [64] * 8
    rows = {}
    for p in plan.pieces:
        assert plan.bounds[p.dst] <= p.s < p.e <= plan.bounds[p.dst + 1]
        rows[p.win] = rows.get(p.win, 0) + (p.e - p.s)
    assert all(v == 128 for v in rows.values()) and len(rows) == 50
    # bytes that cross NVLink (bf16, 10 classes) stay far below the 1.68 GB fp32 accumulator round 1 all-reduced
    moved = sum(v for (s, d), v in plan.pair_voxels.items() if s != d) * 10 * 2
    assert moved < 1.5e9
