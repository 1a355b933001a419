#!/bin/bash
# 8-GPU measurements of round 2 (run under `gpurun --gpus 8`): the scaling bench (cfg3) with its data-parallel proof,
# cfg5 (wide net, 160^3, batch 4/GPU), cfg4 sliding-window inference sharded over 8 and 4 ranks (constant and
# Gaussian importance).  Every JSON line lands in gpurun_out/.
mkdir -p gpurun_out
run() { # name, nproc, port, env..., -- args
  name=$1; np=$2; port=$3; shift 3
  envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  env "${envs[@]}" python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 \
      --master-port $port "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  echo "$name rc=$? $(tail -c 260 gpurun_out/$name.json)"
}
run g8_bench_cfg3 8 29601 X=1 -- bench.py --gpus 8 --steps 50 --warmup 5 --no-roofline
run g8_bench_cfg5 8 29602 X=1 -- bench.py --config cfg5 --gpus 8 --steps 20 --warmup 3 --no-roofline
run g8_infer 8 29603 X=1 -- scripts/infer_bench.py --gpus 8
run g8_infer_gauss 8 29604 X=1 -- scripts/infer_bench.py --gpus 8 --mode gaussian
run g4_infer 4 29605 X=1 -- scripts/infer_bench.py --gpus 4
run g4_bench_cfg3 4 29606 X=1 -- bench.py --gpus 4 --steps 50 --warmup 5 --no-roofline
run g2_bench_cfg3 2 29607 X=1 -- bench.py --gpus 2 --steps 50 --warmup 5 --no-roofline
run g2_infer 2 29608 X=1 -- scripts/infer_bench.py --gpus 2
