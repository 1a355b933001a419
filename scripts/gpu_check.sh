#!/bin/bash
# One gpurun call: GPU parity tests, bench line, per-launch device times of one step.
#   gpurun --timeout 900 -- 'bash scripts/gpu_check.sh [tag] [pytest -k expr]'
tag=${1:-cur}
kexpr=${2:-}
mkdir -p gpurun_out
if [ -n "$kexpr" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/pytest_gpu.log 2>&1
else
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
fi
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
rc=$?
echo "bench exit $rc"; tail -n 3 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.log
if [ $rc -eq 0 ]; then
  timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 700 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline \
    > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu exit $?"
fi
