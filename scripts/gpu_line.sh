#!/bin/bash
# One gpurun call: persistent sliding kernel + line-tiled kernel -- parity tests of the kernels they take over, A/B
# timings, then (only if the tests passed) the whole GPU suite and bench lines for the combinations.
tag=${1:-l2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tcgen05_conv or partials or fused_with" > gpurun_out/${tag}_pytest_k.log 2>&1
rc=$?
echo "kernel pytest exit $rc"; tail -n 25 gpurun_out/${tag}_pytest_k.log
for cfg in "0 0" "0 1" "1 1"; do
  set -- $cfg
  echo "== LINE_CONV=$1 SLIDE_PERSIST=$2" >> gpurun_out/${tag}_head.log
  export B200SEG_LINE_CONV=$1 B200SEG_SLIDE_PERSIST=$2
  timeout 120 python scripts/head_layer.py >> gpurun_out/${tag}_head.log 2>&1
  timeout 120 python scripts/head_layer.py 16 64 >> gpurun_out/${tag}_head.log 2>&1
  timeout 120 python scripts/head_layer.py 32 32 >> gpurun_out/${tag}_head.log 2>&1
  timeout 120 python scripts/fuse_bench.py >> gpurun_out/${tag}_head.log 2>&1
done
cat gpurun_out/${tag}_head.log
if [ $rc -eq 0 ]; then
  export B200SEG_LINE_CONV=1 B200SEG_SLIDE_PERSIST=1
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
  tail -n 8 gpurun_out/${tag}_pytest.log
  for cfg in "0 0" "0 1" "1 1"; do
    set -- $cfg
    B200SEG_LINE_CONV=$1 B200SEG_SLIDE_PERSIST=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/${tag}_bench_$1$2.json 2> gpurun_out/${tag}_bench_$1$2.err
    echo "bench line=$1 persist=$2 exit $?"; tail -n 3 gpurun_out/${tag}_bench_$1$2.err; cut -c1-200 gpurun_out/${tag}_bench_$1$2.json
  done
fi
