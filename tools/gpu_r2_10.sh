#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags lz4,md5,both --iters 2 2> gpurun_out/r2_10_sweep.err > gpurun_out/r2_10_sweep.jsonl
timeout 300 python tools/sweep.py --total-mib 8192 --sizes-mib 1 --workloads random,silesia --flags lz4,md5,both --iters 2 2>> gpurun_out/r2_10_sweep.err >> gpurun_out/r2_10_sweep.jsonl
cat gpurun_out/r2_10_sweep.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_fused -c 1 -o gpurun_out/r2_10_lz4only_random \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads random --flags lz4 --iters 1 > gpurun_out/r2_10_ncu.log 2>&1
echo done
