#!/bin/bash
set +e
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multipart_sized" > gpurun_out/pytest_big.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/pytest_big.log
