#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_12_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_12_parity.log
tail -5 gpurun_out/r2_12_parity.log
timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both --iters 3 2> gpurun_out/r2_12_sweep.err > gpurun_out/r2_12_sweep.jsonl
cat gpurun_out/r2_12_sweep.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_fused -c 1 -o gpurun_out/r2_12_lz4only_silesia \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/r2_12_ncu.log 2>&1
echo done
