#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2_16_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/r2_16_gpu_tests.log
tail -4 gpurun_out/r2_16_gpu_tests.log
for v in pr2_pa10 pr3_pa9 pr4_pa8; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4 --iters 3 \
     2> gpurun_out/r2_16_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> gpurun_out/r2_16_sweep.jsonl
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 4096 --sizes-mib 16 --workloads silesia --flags both --iters 2 \
     2>> gpurun_out/r2_16_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> gpurun_out/r2_16_sweep.jsonl
done
cut -c1-220 gpurun_out/r2_16_sweep.jsonl
timeout 300 python -m skyplane_b200.harness --gpus 1 --chunks 2048 --chunk-mib 8 --pool 32 --workload random > gpurun_out/r2_16_harness_n1.json 2> gpurun_out/r2_16_harness.err
cat gpurun_out/r2_16_harness_n1.json; tail -2 gpurun_out/r2_16_harness.err
echo done
