#!/bin/bash
# One gpurun call: the line-tiled kernel (tc_line.cu) -- parity tests of the kernels it takes over, then A/B timings
# against the sliding kernel, then (only if the tests passed) the whole GPU suite and the bench line.
tag=${1:-l1}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tcgen05_conv or partials or fused_with" > gpurun_out/${tag}_pytest_k.log 2>&1
rc=$?
echo "kernel pytest exit $rc"; tail -n 25 gpurun_out/${tag}_pytest_k.log
for cfg in "0 1" "1 1" "1 2"; do
  set -- $cfg
  echo "== LINE_CONV=$1 EG=$2" >> gpurun_out/${tag}_head.log
  B200SEG_LINE_CONV=$1 B200SEG_LINE_EG=$2 timeout 120 python scripts/head_layer.py >> gpurun_out/${tag}_head.log 2>&1
  B200SEG_LINE_CONV=$1 B200SEG_LINE_EG=$2 timeout 120 python scripts/head_layer.py 16 64 >> gpurun_out/${tag}_head.log 2>&1
  B200SEG_LINE_CONV=$1 B200SEG_LINE_EG=$2 timeout 120 python scripts/fuse_bench.py >> gpurun_out/${tag}_head.log 2>&1
done
cat gpurun_out/${tag}_head.log
if [ $rc -eq 0 ]; then
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -n 8 gpurun_out/${tag}_pytest.log
  for eg in 2 1; do
    B200SEG_LINE_EG=$eg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/${tag}_bench_eg$eg.json 2> gpurun_out/${tag}_bench_eg$eg.err
    echo "bench eg$eg exit $?"; tail -n 3 gpurun_out/${tag}_bench_eg$eg.err; cut -c1-400 gpurun_out/${tag}_bench_eg$eg.json
  done
fi
