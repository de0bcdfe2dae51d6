#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m skyplane_b200.harness --gpus 1 --chunks 4096 --chunk-mib 8 --pool 32 --workload random > gpurun_out/r2_17_harness_n1.json 2> gpurun_out/r2_17_harness.err
cat gpurun_out/r2_17_harness_n1.json; tail -2 gpurun_out/r2_17_harness.err
timeout 300 python -m skyplane_b200.harness --gpus 1 --chunks 4096 --chunk-mib 8 --pool 32 --workload random --batch 64 --slots 4 > gpurun_out/r2_17_harness_n1_b64s4.json 2>> gpurun_out/r2_17_harness.err
cat gpurun_out/r2_17_harness_n1_b64s4.json
timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,md5,both --iters 3 --decode 2> gpurun_out/r2_17_sweep.err > gpurun_out/r2_17_sweep.jsonl
cut -c1-200 gpurun_out/r2_17_sweep.jsonl
echo done
