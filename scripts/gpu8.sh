#!/bin/bash
# 8-GPU measurements of round 2 (run under `gpurun --gpus 8`): scaling bench with NCCL CTA caps, cfg4 sliding-window
# inference sharded over 8 ranks, cfg5 (wide net, 160^3, batch 4/GPU).  Every JSON line lands in gpurun_out/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run() { # name, port, env..., -- args
  name=$1; port=$2; shift 2
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" $TR --master-port $port "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(tail -c 300 gpurun_out/$name.json | head -c 300)"
}
run g8_bench_default 29601 X=1 -- bench.py --gpus 8 --steps 30 --warmup 5 --no-roofline
run g8_bench_cta8 29602 NCCL_MAX_CTAS=8 -- bench.py --gpus 8 --steps 30 --warmup 5 --no-roofline
run g8_bench_cta4 29603 NCCL_MAX_CTAS=4 -- bench.py --gpus 8 --steps 30 --warmup 5 --no-roofline
run g8_bench_cta16 29604 NCCL_MAX_CTAS=16 -- bench.py --gpus 8 --steps 30 --warmup 5 --no-roofline
run g8_infer 29605 X=1 -- scripts/infer_bench.py --gpus 8
run g8_infer_gauss 29606 X=1 -- scripts/infer_bench.py --gpus 8 --mode gaussian
run g8_cfg5 29607 X=1 -- bench.py --config cfg5 --gpus 8 --steps 10 --warmup 3 --no-roofline
