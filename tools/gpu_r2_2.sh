#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in e4096_w13 e3072_w18 e2048_w26 e2048_w16; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both --iters 3 \
     2> gpurun_out/r2_2_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> gpurun_out/r2_2_sweep.jsonl
done
cat gpurun_out/r2_2_sweep.jsonl
