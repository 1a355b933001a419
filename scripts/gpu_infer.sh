#!/bin/bash
# cfg4 on one B200: inference parity tests, then scripts/infer_bench.py (512x512x160 volume, 50 windows)
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_inference.py -m gpu -x -q > gpurun_out/pytest_infer.log 2>&1
echo "pytest exit $?"; tail -n 12 gpurun_out/pytest_infer.log
timeout 120 python scripts/infer_bench.py > gpurun_out/infer_bench.log 2> gpurun_out/infer_bench.err
echo "infer exit $?"; cat gpurun_out/infer_bench.log; tail -n 5 gpurun_out/infer_bench.err
