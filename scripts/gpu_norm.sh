#!/bin/bash
# One gpurun call: cluster InstanceNorm backward + final slide / line dispatch: tests, then bench A/B.
tag=${1:-n1}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "instnorm or tcgen05_conv or partials or fused_with" > gpurun_out/${tag}_pytest_k.log 2>&1
rc=$?
echo "kernel pytest exit $rc"; tail -n 25 gpurun_out/${tag}_pytest_k.log
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -n 8 gpurun_out/${tag}_pytest.log
  for cfg in "0 0" "0 1" "1 1" "1 0"; do
    set -- $cfg
    B200SEG_LINE_CONV=$1 B200SEG_NORM_CLUSTER=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/${tag}_bench_$1$2.json 2> gpurun_out/${tag}_bench_$1$2.err
    echo "bench line=$1 cluster=$2 exit $?"; tail -n 3 gpurun_out/${tag}_bench_$1$2.err; cut -c1-200 gpurun_out/${tag}_bench_$1$2.json
  done
fi
