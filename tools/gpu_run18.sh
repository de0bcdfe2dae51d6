#!/bin/bash
# bounded validation after the pacing-overflow fix (every command has a short timeout)
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 240 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== sweep silesia 1MiB"; timeout 90 python tools/sweep.py --total-mib 2048 --sizes-mib 1,8 --workloads silesia --flags lz4,both --iters 2 > gpurun_out/sweep_fix.jsonl 2> gpurun_out/sweep_fix.err; echo "rc=$?"; cut -c1-200 gpurun_out/sweep_fix.jsonl; tail -2 gpurun_out/sweep_fix.err
echo "== racecheck"; timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or mixed" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck.log
