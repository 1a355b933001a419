#!/bin/bash
# N-GPU box (gpurun --gpus 2): gradient-exchange check of GraphedTrainStep, then the 2-GPU bench line
mkdir -p gpurun_out
timeout 150 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/ddp_overlap_check.py > gpurun_out/ddp_check.log 2> gpurun_out/ddp_check.err
echo "ddp check exit $?"; cat gpurun_out/ddp_check.log; grep -v "^\*\|OMP_NUM\|\.\.\." gpurun_out/ddp_check.err | tail -n 15
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/bench_2gpu_r1end.log 2> gpurun_out/bench_2gpu_r1end.err
echo "bench2 exit $? $(python -c "import json; d=json.loads(open('gpurun_out/bench_2gpu_r1end.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['loss'])")"
