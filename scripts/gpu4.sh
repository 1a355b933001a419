#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus 4 --steps 40 --warmup 5 --no-roofline > gpurun_out/bench_4gpu_r1end.log 2> gpurun_out/bench_4gpu_r1end.err
echo "bench4 exit $? $(python -c "import json; d=json.loads(open('gpurun_out/bench_4gpu_r1end.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['loss'])")"
grep -v "^\*\|OMP_NUM" gpurun_out/bench_4gpu_r1end.err | tail -n 8
