#!/bin/bash
# Line kernel with the one-voxel-per-thread epilogue: parity tests with it switched on, timings, bench A/B.
tag=${1:-l3}
mkdir -p gpurun_out
export B200SEG_LINE_CONV=1 B200SEG_LINE_W128=1
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tcgen05_conv or partials or fused_with" > gpurun_out/${tag}_pytest_k.log 2>&1
rc=$?
echo "kernel pytest (line on) exit $rc"; tail -n 15 gpurun_out/${tag}_pytest_k.log
for dbg in 0 1; do
  echo "== LINE on, DEBUG=$dbg" >> gpurun_out/${tag}_head.log
  B200SEG_LINE_DEBUG=$dbg timeout 60 python scripts/head_layer.py 2>&1 | grep -E "fprop|dgrad" >> gpurun_out/${tag}_head.log
done
timeout 120 python scripts/head_layer.py 16 64 2>&1 | grep -E "fprop|dgrad" >> gpurun_out/${tag}_head.log
timeout 120 python scripts/fuse_bench.py >> gpurun_out/${tag}_head.log 2>&1
echo "== LINE off" >> gpurun_out/${tag}_head.log
B200SEG_LINE_CONV=0 timeout 120 python scripts/fuse_bench.py >> gpurun_out/${tag}_head.log 2>&1
cat gpurun_out/${tag}_head.log
if [ $rc -eq 0 ]; then
  timeout 300 python -m pytest tests/test_gpu_unet.py tests/test_gpu_parity_shapes.py tests/test_gpu_guards.py -m gpu -x -q > gpurun_out/${tag}_pytest_net.log 2>&1
  echo "net pytest (line on) exit $?"; tail -n 3 gpurun_out/${tag}_pytest_net.log
  for cfg in "1 1" "0 0" "1 0"; do
    set -- $cfg
    B200SEG_LINE_CONV=$1 B200SEG_LINE_W128=$2 timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/${tag}_bench_$1$2.json 2> gpurun_out/${tag}_bench_$1$2.err
    echo "bench line=$1 w128=$2 exit $?"; tail -n 2 gpurun_out/${tag}_bench_$1$2.err; cut -c1-200 gpurun_out/${tag}_bench_$1$2.json
  done
fi
