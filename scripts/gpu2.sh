#!/bin/bash
# 2-GPU box: engine test, gradient-exchange check, bench with / without the overlapped all-reduce
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "graphed_train_step or overlapped_allreduce" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit $?"; tail -n 15 gpurun_out/pytest_gpu2.log
for ov in 1 0; do
  B200SEG_OVERLAP_ALLREDUCE=$ov timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/bench_2gpu_ov$ov.log 2> gpurun_out/bench_2gpu_ov$ov.err
  echo "overlap=$ov exit $? $(python -c "import json; d=json.loads(open('gpurun_out/bench_2gpu_ov$ov.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'], d['loss'])")"
  tail -n 3 gpurun_out/bench_2gpu_ov$ov.err
done
timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_1gpu_b.log 2> gpurun_out/bench_1gpu_b.err
echo "1gpu exit $? $(python -c "import json; d=json.loads(open('gpurun_out/bench_1gpu_b.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['value'])")"
