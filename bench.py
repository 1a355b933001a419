#!/usr/bin/env python
"""Headline benchmark: 3D U-Net train voxels/s (fwd + Dice loss + Dice metric + bwd [+ grad all-reduce] + Adam).

    python bench.py --gpus N --steps K --warmup W            # this framework (B200 kernels)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle)
    python bench.py --config cfg2|cfg3|cfg5 ...              # BASELINE.json configs[1] / [2] (default) / [4]

Workloads (BASELINE.json `configs`): MONAI residual U-Net (in=1, out=10, num_res_units=2), synthetic patches, bf16
storage / fp32 accumulation, softmax Dice loss, weak scaling (fixed patches per GPU), data-parallel over N GPUs
with one flat NCCL all-reduce of the gradients per step, overlapped with the backward pass inside the CUDA graph.
    cfg2 = configs[1]: 16-32-64-128-256, 96^3,  batch 2, single B200
    cfg3 = configs[2]: 16-32-64-128-256, 128^3, batch 2 / GPU, 1/2/4/8 B200       (the headline)
    cfg5 = configs[4]: 32-64-128-256-512, 160^3, batch 4 / GPU, 8 x B200
One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FWD_BWD_FLOP_PER_VOXEL = {(16, 32, 64, 128, 256): 55236.0, (32, 64, 128, 256, 512): 158520.0}
CONFIGS = {  # name -> (BASELINE.json configs index, filters, patch, batch per GPU)
    "cfg2": (1, [16, 32, 64, 128, 256], 96, 2),
    "cfg3": (2, [16, 32, 64, 128, 256], 128, 2),
    "cfg5": (4, [32, 64, 128, 256, 512], 160, 4),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg3", choices=sorted(CONFIGS))
    ap.add_argument("--patch", type=int, default=None, help="override the config's patch edge")
    ap.add_argument("--batch", type=int, default=None, help="override the config's patches per GPU per step")
    ap.add_argument("--filters", type=int, nargs=5, default=None, help="override the config's channel widths")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-graph", action="store_true", help="issue every launch eagerly (no CUDA graph)")
    ap.add_argument("--no-metric", action="store_true", help="leave the per-step Dice-metric pass out of the step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--no-gpu-bar", action="store_true")
    args = ap.parse_args()
    idx, filters, patch, batch = CONFIGS[args.config]
    args.config_index = idx
    args.filters = args.filters or filters
    args.patch = args.patch or patch
    args.batch = args.batch or batch
    return args


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"],
                "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


def synthetic_batch(batch, patch, seed):
    """Seeded synthetic inputs: randn images, dense-random labels 0..9."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 1, patch, patch, patch, generator=g)
    labels = torch.randint(0, 10, (batch, patch, patch, patch), generator=g, dtype=torch.uint8)
    return images, labels


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""

    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def mark(self):
        """Start of the timed region: only samples taken after this call are reported."""
        self.t0 = time.time()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        t1 = time.time()
        time.sleep(0.1)
        self.proc.terminate()
        t0 = getattr(self, "t0", 0.0)
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1] or [r for _, r in self.rows[-3:]]
        self.rows = rows
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def cpu_reference_step_time(filters, patch, batch, steps, warmup, threads, optimizer=False, budget_s=None):
    """The reference's path (oracle restatement: MONAI-0.3 UNet + DiceLoss on torch CPU, fp32, [+ Adam]).
    ``budget_s``: stop timing early once the timed steps have used that many seconds (bounded sample)."""
    from oracle import monai_ref as O
    torch.set_num_threads(threads)
    torch.manual_seed(12342)
    net = O.UNet(3, 1, 10, filters, [2, 2, 2, 2], num_res_units=2)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3) if optimizer else None
    images, labels = synthetic_batch(batch, patch, 12342)
    labels = labels.long()
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(net, images, labels)
        if opt is not None:
            opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
            if budget_s is not None and sum(times) > budget_s:
                break
    return times


def run_reference(args, rank, world):
    """`--impl reference`: the reference's own CPU implementation of the path (oracle port), on the SAME workload as
    the B200 arm -- same network, patch, per-GPU batch, Dice loss and Adam, every host thread.  Bounded: the timed
    steps stop after ~100 s of CPU work (cfg5's step alone takes about a minute); `steps` reports what ran."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    warmup = max(1, args.warmup) if args.config != "cfg5" else 1
    times = cpu_reference_step_time(args.filters, args.patch, args.batch, max(1, args.steps), warmup, threads,
                                    optimizer=True, budget_s=100.0)
    steps = len(times)
    mean_t = sum(times) / steps
    vox = args.batch * args.patch ** 3
    value = vox / mean_t
    sample = (f"{steps} timed steps (of {args.steps} requested; ~100 s budget) of {args.batch}x{args.patch}^3 "
              f"fwd+Dice+bwd+Adam after {warmup} warm-up, torch CPU fp32, {threads} threads, one process")
    print(json.dumps({
        "impl": "reference", "metric": "3D U-Net train voxels/sec (fwd+bwd)", "value": value,
        "unit": "voxels/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": mean_t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, args.gpus, reference=True),
        "cpu_baseline": {"value": value, "unit": "voxels/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, world, reference=False):
    cfg = {
        "workload": f"3D MONAI residual UNet {'-'.join(map(str, args.filters))} (in=1,out=10,res_units=2), "
                    f"{args.patch}^3 patches, batch {args.batch}/GPU, softmax Dice loss, "
                    f"fwd+bwd+Adam (BASELINE.json configs[{args.config_index}])",
        "name": args.config, "patch": args.patch, "batch_per_gpu": args.batch, "global_batch": args.batch * world,
        "parallelism": f"dp{world}",
    }
    if reference:
        cfg["optimizer"] = "torch.optim.Adam, inside the timed region"
        cfg["execution"] = ("reference CPU path: oracle port (MONAI-0.3 UNet + DiceLoss restated on torch.nn, fp32) on "
                            "the host cores of rank 0; one process (the reference has no CPU data parallelism), so the "
                            "per-step workload is ONE rank's share of the global batch")
        return cfg
    cfg["optimizer"] = "Adam (FlatAdam: torch.optim.Adam semantics, one launch; inside the timed region)"
    cfg["execution"] = "eager launches" if getattr(args, "no_graph", False) else (
        "fwd+loss+bwd replayed from one CUDA graph" +
        (" incl. the NCCL gradient all-reduce (deep levels' range overlapped with the rest of the backward pass)"
         if world > 1 else ""))
    cfg["dice_metric"] = ("left out (--no-metric)" if getattr(args, "no_metric", False) else
                          "per-step Dice metric (reference _log_dice_scores) inside the step: argmax + integer counts in the "
                          "Dice loss's own pass over the logits, the (B, 9) epilogue on a graph branch")
    cfg["l2"] = "per-step activations+gradients (>1 GB) exceed the 126 MB L2; no explicit flush"
    return cfg


# ----------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local):
    import torch.distributed as dist

    import ct_image_segmentation_b200 as B
    from ct_image_segmentation_b200 import _lib, ops
    from ct_image_segmentation_b200.parallel import GradientBucket

    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    _lib.check(lib.b200seg_check_device(local), "b200seg_check_device")
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(12342)
    net = B.UNet(3, 1, 10, args.filters, [2, 2, 2, 2], num_res_units=2, dtype=dtype).to(dev)
    loss_fx = B.DiceLoss(include_background=False, to_onehot_y=True, softmax=True, with_metric=not args.no_metric)
    opt = B.FlatAdam(net.parameters(), lr=1e-3)  # torch.optim.Adam semantics, one launch on the flat bucket
    images_h, labels_h = synthetic_batch(args.batch, args.patch, 12342 + rank)
    images_h, labels_h = images_h.pin_memory(), labels_h.pin_memory()
    images_d, labels_d = images_h.to(dev), labels_h.to(dev)
    vox_per_step = args.batch * args.patch ** 3 * world
    # public API: forward + Dice loss + Dice metric + backward (+ gradient exchange) as one CUDA graph, then Adam
    metric = None if args.no_metric else "fused"
    train = B.GraphedTrainStep(net, loss_fx, opt, images_d, labels_d, use_graph=not args.no_graph, metric=metric)

    def step(images, labels):
        return train(images, labels)

    def timed(n_steps, from_host):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = pending = None
        for _ in range(n_steps):
            if from_host:
                # every step: H2D of its inputs (pinned host memory, copy stream) and a D2H read of a
                # step result; the read is of the PREVIOUS step's loss so that the host can enqueue
                # step k+1's transfer while step k computes (software pipelining, nothing skipped)
                loss = step(images_h, labels_h)
                host_loss = torch.empty((), dtype=loss.dtype, pin_memory=True)
                host_loss.copy_(loss, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[1].synchronize()
                    last = pending[0].item()
                pending = (host_loss, ev)
            else:
                last = step(None, None)  # inputs already resident in the graph's input buffers
        if from_host and pending is not None:
            pending[1].synchronize()
            last = pending[0].item()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, (last if from_host else last.item())

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    timed(args.warmup, False)
    if sampler:
        sampler.mark()
    l0 = lib.b200seg_launch_count()
    ms_dev, loss_val = timed(args.steps, False)
    launches = lib.b200seg_launch_count() - l0
    if train.launches_per_step is not None:  # graph replay: the recorded kernels run once per step
        launches = train.launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None
    timed(1, True)
    ms_e2e, _ = timed(args.steps, True)

    value = vox_per_step * args.steps / (ms_dev * 1e-3)
    e2e = vox_per_step * args.steps / (ms_e2e * 1e-3)
    pk = peaks()
    out = {
        "metric": "3D U-Net train voxels/sec (fwd+bwd)", "value": value, "unit": "voxels/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e, "unit": "voxels/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": images_h.numel() * 4 + labels_h.numel(), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "tcgen05_launches_per_step": train.tc_launches_per_step,
        "loss": loss_val, "clocks": clocks,
    }
    flop = FWD_BWD_FLOP_PER_VOXEL.get(tuple(args.filters))
    if flop:
        out["step_tflops"] = flop * vox_per_step / world * args.steps / (ms_dev * 1e-3) / 1e12
    if metric is not None and train.metric_out is not None:
        out["dice_metric_mean"] = float(train.metric_out[0])
    if world > 1:
        # data-parallel proof (outside the timed region): the bucket after the in-graph overlapped exchange is the
        # mean over ranks of the ranks' local gradients (one plain all-reduce of separately computed local gradients)
        chk = train.check_exchange()
        t = torch.tensor([chk["max_rel_to_largest"]], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        chk["max_rel_to_largest"] = t.item()
        chk["ok"] = chk["max_rel_to_largest"] < 1e-5
        out["ddp_check"] = chk
    # free the training graph's memory before the side measurements
    del train
    torch.cuda.empty_cache()
    if rank == 0 and not args.no_roofline:
        fams = roofline_families(args, dev, dtype, pk)
        out["roofline"] = fams[0]
        out["roofline_families"] = fams[1:]
    if rank == 0 and world == 1 and not args.no_gpu_bar:
        out["gpu_bar"] = gpu_bar(args, dev, out["ms_per_step"])
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_steps = 3 if args.config != "cfg5" else 1
        times = cpu_reference_step_time(args.filters, args.patch, 1, n_steps, 1, threads)
        best = min(times)
        out["cpu_baseline"] = {
            "value": args.patch ** 3 / best, "unit": "voxels/s", "cores": threads, "kind": "port",
            "sample": f"best of {n_steps} steps of 1x{args.patch}^3 fwd+Dice+bwd on torch CPU fp32 after 1 warm-up",
        }
    if rank == 0:
        print(json.dumps(out))
    if world > 1 and rank == 0 and not out["ddp_check"]["ok"]:
        raise SystemExit("data-parallel check failed: bucket != mean of local gradients")


def _graph_time_us(fn, reps=10, warm=3):
    """us per call of `fn` (launches on the current stream), replayed `reps` times per CUDA-graph launch so that no
    host launch latency is in the number; CUDA events on the launching stream around 3 replays."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warm):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3


def _ncu_traffic(key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
    (profiles/ncu_traffic.json: {key: {"bytes": ..., "source": "<profile file>"}}); None when not captured."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        return tab.get(key)
    except Exception:  # noqa: BLE001
        return None


def roofline_families(args, dev, dtype, pk):
    """The kernel families that carry the step, each timed alone at the configuration's own shapes (CUDA events on the
    launching stream, graph-replayed launches) against the roofline that bounds it:
    HBM for the 10/16-channel full-resolution layers and the InstanceNorm / loss passes (algorithmic bytes = every
    operand read once + every result written once, UNPADDED channel counts), tensor peak (burst, kernel timed alone)
    for the deep C >= 64 convolutions.  Entry 0 is the dominant kernel (largest share of the step)."""
    from ct_image_segmentation_b200 import _lib, ops
    lib = _lib.load()
    n, p = args.batch, args.patch
    c0 = args.filters[0]
    esz = 2 if dtype == torch.bfloat16 else 4
    hbm, tc = pk["hbm_gbs"], pk["bf16_tflops"]
    fams = []

    def rand_act(nb, sp, c):
        t = ops.alloc_activation(nb, sp, c, dtype, dev)
        t.copy_(torch.randn(t.shape, device=dev))
        return t

    def entry(name, kernel, us, bytes_alg, flops, bound, key=None, note=None):
        gbs, tfs = bytes_alg / (us * 1e-6) / 1e9, flops / (us * 1e-6) / 1e12
        e = {"name": name, "kernel": kernel, "bound": bound, "us": us,
             "achieved": gbs if bound == "hbm" else tfs, "peak": hbm if bound == "hbm" else tc,
             "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
             "frac": (gbs / hbm) if bound == "hbm" else (tfs / tc),
             "algorithmic_bytes": bytes_alg, "algorithmic_flops": flops, "achieved_gbs": gbs, "achieved_tflops": tfs,
             "peak_source": pk["src"] + (" (HBM copy)" if bound == "hbm" else " (bf16 burst: kernel timed alone)")}
        tr = _ncu_traffic(key) if key else None
        e["traffic"] = tr["bytes"] if tr else None
        if tr:
            e["traffic_source"] = tr.get("source")
        if note:
            e["note"] = note
        fams.append(e)

    def conv_layer(name, cin, cout, k, s, tr, nb, sp_in, bound, what=("fprop", "dgrad", "wgrad"), key=None):
        g = ops.ConvGeom(3, cin, cout, k, s, tr)
        sp_out = g.out_spatial(*sp_in)
        x, dy = rand_act(nb, sp_in, cin), rand_act(nb, sp_out, cout)
        y, dx = ops.alloc_activation(nb, sp_out, cout, dtype, dev), ops.alloc_like(x)
        w = torch.randn((cin, cout, k, k, k) if tr else (cout, cin, k, k, k), device=dev) * 0.05
        wf = ops.pack_weight(g, _lib.W_CONVTR_FPROP if tr else _lib.W_CONV_FPROP, w, dtype)
        wd = ops.pack_weight(g, _lib.W_CONVTR_DGRAD if tr else _lib.W_CONV_DGRAD, w, dtype)
        b = torch.zeros(cout, device=dev)
        vi, vo = nb * sp_in[0] * sp_in[1] * sp_in[2], nb * sp_out[0] * sp_out[1] * sp_out[2]
        macs = (vi if tr else vo) * cin * cout * k ** 3
        bytes_io = (vi * cin + vo * cout) * esz
        runs = {"fprop": lambda: ops.conv_fprop(g, x, wf, b, y), "dgrad": lambda: ops.conv_dgrad(g, dy, wd, dx),
                "wgrad": lambda: ops.conv_wgrad(g, x, dy, want_bias=False)}
        for op in what:
            us = _graph_time_us(runs[op])
            kern = lib.b200seg_last_launch().decode()
            entry(f"{name} {op}", kern, us, bytes_io, 2.0 * macs, bound, key=f"{key}:{op}" if key else None)

    # (0) dominant: head conv 10->10 k3 s1 at full resolution (29 % of the conv FLOPs; AI 135 < ridge 211: HBM)
    conv_layer(f"head conv 10->10 k3 s1 @{p}^3 x{n}", 10, 10, 3, 1, False, n, (p, p, p), "hbm",
               key=f"head_{p}_{n}")
    # ConvTranspose c1->10 to full resolution
    c1 = 2 * c0
    conv_layer(f"ConvT {c1}->10 k3 s2 @{p // 2}^3->{p}^3 x{n}", c1, 10, 3, 2, True, n, (p // 2,) * 3, "hbm",
               key=f"convtr_{p}_{n}")
    # level-0 residual-unit conv c0->c0 at half resolution (AI ~ ridge)
    conv_layer(f"conv {c0}->{c0} k3 s1 @{p // 2}^3 x{n}", c0, c0, 3, 1, False, n, (p // 2,) * 3,
               "hbm" if c0 <= 16 else "tensor")
    # deep levels: tensor-bound by arithmetic intensity
    c3, c4 = args.filters[3], args.filters[4]
    conv_layer(f"conv {args.filters[2]}->{args.filters[2]} k3 s1 @{p // 8}^3 x{n}", args.filters[2], args.filters[2], 3, 1,
               False, n, (p // 8,) * 3, "tensor")
    conv_layer(f"conv {c4}->{c4} k3 s1 @{p // 16}^3 x{n} (bottom)", c4, c4, 3, 1, False, n, (p // 16,) * 3, "tensor")
    # InstanceNorm + PReLU passes on the full-resolution 10-class tensor and on the c0 tensor at half resolution
    for cc, sp, tag in ((10, (p, p, p), f"10ch @{p}^3"), (c0, (p // 2,) * 3, f"{c0}ch @{p // 2}^3")):
        x, dy = rand_act(n, sp, cc), rand_act(n, sp, cc)
        y, dx = ops.alloc_like(x), ops.alloc_like(x)
        alpha = torch.full((1,), 0.25, device=dev)
        mean, rstd = ops.instnorm_stats(x)
        elems = n * sp[0] * sp[1] * sp[2] * cc
        us = _graph_time_us(lambda: ops.instnorm_prelu_fwd(x, mean, rstd, alpha, y))
        entry(f"InstanceNorm+PReLU fwd apply {tag} x{n}", lib.b200seg_last_launch().decode(), us, 2 * esz * elems, 0,
              "hbm", key=f"in_fwd_{cc}_{p}_{n}")
        us = _graph_time_us(lambda: ops.instnorm_prelu_bwd(x, mean, rstd, alpha, dy, dx))
        entry(f"InstanceNorm+PReLU bwd (reduce + apply) {tag} x{n}", lib.b200seg_last_launch().decode(), us,
              5 * esz * elems, 0, "hbm", key=f"in_bwd_{cc}_{p}_{n}",
              note="two passes: reductions (reads dy, x) then dx (reads dy, x, writes dx) = 5 e per element")
    # the pair the step actually runs for the head layer and the c0 = 16 layers: dgrad with the InstanceNorm-backward
    # sums in its epilogue, then final reduction + apply
    if dtype == torch.bfloat16:
        for cc, sp, tag in ((10, (p, p, p), f"head 10->10 @{p}^3"), (c0, (p // 2,) * 3, f"{c0}->{c0} @{p // 2}^3")):
            g = ops.ConvGeom(3, cc, cc, 3, 1, False)
            cprev, dy, res = rand_act(n, sp, cc), rand_act(n, sp, cc), rand_act(n, sp, cc)
            dx, gc = ops.alloc_like(cprev), ops.alloc_like(cprev)
            wd = ops.pack_weight(g, _lib.W_CONV_DGRAD, torch.randn(cc, cc, 3, 3, 3, device=dev) * 0.1, dtype)
            mean, rstd = ops.instnorm_stats(cprev)
            alpha = torch.full((1,), 0.25, device=dev)
            h = ops.conv_dgrad_instnorm_partials(g, dy, wd, dx, cprev, mean, rstd, alpha, residual=res)
            if h is None:
                continue
            vox_l = n * sp[0] * sp[1] * sp[2]
            us = _graph_time_us(lambda: ops.conv_dgrad_instnorm_partials(g, dy, wd, dx, cprev, mean, rstd, alpha,
                                                                         residual=res))
            entry(f"{tag} x{n} dgrad + residual + InstanceNorm-backward sums (fused epilogue)",
                  lib.b200seg_last_launch().decode(), us, vox_l * cc * 4 * esz, 2.0 * 27 * cc * cc * vox_l, "hbm",
                  note="reads dy, the residual addend and the consumer layer's pre-norm tensor, writes dx: 4 e per element")
            us = _graph_time_us(lambda: ops.instnorm_prelu_bwd_from_partials(cprev, mean, rstd, alpha, dx, gc, h))
            entry(f"InstanceNorm+PReLU bwd after the fused sums (final + apply) {cc}ch x{n}",
                  "instnorm_bwd_final + instnorm_prelu_bwd_apply", us, vox_l * cc * 3 * esz, 0, "hbm")
    # softmax + Dice, forward sums and backward
    z = rand_act(n, (p, p, p), 10)
    lab = torch.randint(0, 10, (n, p, p, p), device=dev, dtype=torch.uint8)
    vox = n * p ** 3
    us = _graph_time_us(lambda: ops.softmax_dice_metric_sums(z, lab))
    entry(f"softmax+Dice fwd incl. Dice-metric counts @{p}^3 x{n}", "softmax_dice_fwd_ring + dice_metric_final", us,
          vox * (10 * esz + 1), 0, "hbm", key=f"dice_fwd_{p}_{n}")
    gi = torch.rand(n, 10, device=dev)
    dz = ops.alloc_like(z)  # (the output buffer is the caller's: no allocation / zero fill inside the timed launches)
    us = _graph_time_us(lambda: ops.softmax_dice_bwd(z, lab, gi, gi, dlogits=dz))
    entry(f"softmax+Dice bwd @{p}^3 x{n}", "softmax_dice_bwd_ring", us, vox * (20 * esz + 1), 0, "hbm", key=f"dice_bwd_{p}_{n}")
    fams[0]["note"] = ("dominant kernel: head conv fprop (its dgrad is the same kernel with mirrored taps); HBM is the bound "
                       "by arithmetic intensity (135 FLOP/B < ridge 211); algorithmic bytes = 10 + 10 channels x 2 B per "
                       "voxel (unpadded)")
    return fams


def gpu_bar(args, dev, ours_ms):
    """Same-box GPU bar: the reference's path as it runs on this GPU today -- the oracle's
    MONAI-0.3 UNet + DiceLoss + Adam on stock torch/cuDNN kernels, (i) as the reference would run it (fp32 NCDHW, TF32
    convolutions) and (ii) its strongest fair variant (bf16 autocast + channels_last_3d) -- same workload, same run.
    A baseline leg like `cpu_baseline`: the oracle is only ever the thing compared against."""
    from oracle import monai_ref as O
    res = {"workload": f"oracle UNet + DiceLoss + Adam, {args.batch}x{args.patch}^3, torch {torch.__version__} eager, "
                       f"cudnn.benchmark"}
    old = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(12342)
    images = torch.randn(args.batch, 1, args.patch, args.patch, args.patch, device=dev)
    labels = torch.randint(0, 10, (args.batch, args.patch, args.patch, args.patch), device=dev)
    loss_fx = O.DiceLoss(include_background=False, to_onehot_y=True, softmax=True)
    try:
        for tag, autocast, cl in (("fp32_tf32_ncdhw", False, False), ("bf16_autocast_channels_last_3d", True, True)):
            try:
                net = O.UNet(3, 1, 10, args.filters, [2, 2, 2, 2], num_res_units=2).to(dev)
                x = images
                if cl:
                    net = net.to(memory_format=torch.channels_last_3d)
                    x = images.contiguous(memory_format=torch.channels_last_3d)
                opt = torch.optim.Adam(net.parameters(), lr=1e-3)
                times = []
                for i in range(2 + 5):
                    torch.cuda.synchronize()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    opt.zero_grad(set_to_none=True)
                    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                        y = net(x)
                    loss_fx(y.float(), labels.unsqueeze(1)).backward()
                    opt.step()
                    e1.record()
                    torch.cuda.synchronize()
                    if i >= 2:
                        times.append(e0.elapsed_time(e1))
                res[tag] = {"ms_per_step": min(times), "speedup_of_this_path": min(times) / ours_ms}
                del net, opt
            except Exception as e:  # noqa: BLE001 -- report and go on with the next variant
                res[tag] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark = old
    return res


def main():
    args = parse()
    from ct_image_segmentation_b200.parallel import init_distributed
    if args.impl == "reference":
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        run_reference(args, rank, world)
        return
    rank, world, local = init_distributed()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    run_b200(args, rank, world, local)
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
