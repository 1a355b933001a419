#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -u -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  scripts/ddp_overlap_check.py > gpurun_out/ddp_check.log 2> gpurun_out/ddp_check.err
echo "ddp check exit $?"; cat gpurun_out/ddp_check.log; grep -v "^\*\|OMP_NUM" gpurun_out/ddp_check.err | tail -n 25
