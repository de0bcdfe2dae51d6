#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== sweep"; timeout 900 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads random,silesia --flags md5,both --decode > gpurun_out/sweep_v6.jsonl 2> gpurun_out/sweep_v6.err; cut -c1-200 gpurun_out/sweep_v6.jsonl; tail -3 gpurun_out/sweep_v6.err
