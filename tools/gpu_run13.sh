#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== sanitizer decode"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_decode.py -m gpu -x -q -k "malformed or matches_reference" > gpurun_out/memcheck_decode.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_decode.log
