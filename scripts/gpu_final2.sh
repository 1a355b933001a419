#!/bin/bash
# Round-2 (second session) validation on one B200: GPU parity tests (default dispatch, then once more with the
# experimental kernels switched on), smoke, default bench line (roofline + cpu baseline), A/B of the dalpha fold,
# per-launch device times of one step.
tag=${1:-f1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
tail -n 6 gpurun_out/${tag}_pytest.log
B200SEG_LINE_CONV=1 B200SEG_LINE_W128=1 timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_unet.py -m gpu -x -q > gpurun_out/${tag}_pytest_line.log 2>&1
echo "pytest (line kernel on) exit $?" >> gpurun_out/${tag}_pytest_line.log
tail -n 3 gpurun_out/${tag}_pytest_line.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
echo "smoke exit $?"; tail -n 3 gpurun_out/${tag}_smoke.log
timeout 400 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
rc=$?
echo "bench exit $rc"; tail -n 3 gpurun_out/${tag}_bench.err; cut -c1-600 gpurun_out/${tag}_bench.json
B200SEG_NORM_CLUSTER=0 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/${tag}_bench_nocluster.json 2> gpurun_out/${tag}_bench_nocluster.err
cut -c1-200 gpurun_out/${tag}_bench_nocluster.json
if [ $rc -eq 0 ]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv \
    --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline \
    > gpurun_out/${tag}_ncu.log 2>&1
  echo "ncu exit $?"
fi
