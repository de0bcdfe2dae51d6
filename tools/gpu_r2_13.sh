#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_13_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2_13_gpu_tests.log
tail -30 gpurun_out/r2_13_gpu_tests.log
echo done
