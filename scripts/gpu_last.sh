#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -n 5 gpurun_out/pytest_gpu.log
timeout 100 python -m tests.gpu_bar > gpurun_out/gpu_bar.log 2> gpurun_out/gpu_bar.err
echo "gpu_bar exit $?"; cat gpurun_out/gpu_bar.log; tail -n 3 gpurun_out/gpu_bar.err
