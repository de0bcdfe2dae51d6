#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_operator.py tests/test_gpu_decode.py -m gpu -x -q -k "operator" > gpurun_out/pytest_ops.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_ops.log
