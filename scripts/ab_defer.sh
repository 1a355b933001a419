#!/bin/bash
# A/B of B200SEG_DEFER_STATS_MAX_VOX (InstanceNorm+PReLU finalising the conv's partial statistics itself)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 8 gpurun_out/pytest_gpu.log
for v in 0 4096 32768 262144 99999999999; do
  B200SEG_DEFER_STATS_MAX_VOX=$v timeout 200 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-roofline \
    > gpurun_out/bench_defer_$v.log 2> gpurun_out/bench_defer_$v.err
  echo "defer<=$v exit $? $(python -c "import json,sys; d=json.loads(open('gpurun_out/bench_defer_$v.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'])")"
done
