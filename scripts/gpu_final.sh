#!/bin/bash
# Round-end validation on one B200: GPU parity tests, smoke, default bench line, per-launch device times of one step.
tag=${1:-final}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 6 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_$tag.log 2> gpurun_out/bench_$tag.err
rc=$?
echo "bench exit $rc"; tail -n 3 gpurun_out/bench_$tag.err; cat gpurun_out/bench_$tag.log
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline \
    > gpurun_out/ncu_$tag.log 2>&1
  echo "ncu exit $?"
fi
