#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 12 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/bench_sink.log 2> gpurun_out/bench_sink.err
echo "bench exit $?"; tail -n 3 gpurun_out/bench_sink.err; python -c "import json; d=json.loads(open('gpurun_out/bench_sink.log').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['loss'], d['gpu_launches'])"
