#!/bin/bash
# round 2, GPU call 1: parity of the new compressor + first speed numbers for the tuning builds
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2_1_gpu.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_1_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_1_parity.log
tail -15 gpurun_out/r2_1_parity.log
for v in e4096_w13 e4096_w12 e3072_w18 e2048_w26 e2048_w16; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both --iters 3 \
     2> gpurun_out/r2_1_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> gpurun_out/r2_1_sweep.jsonl
done
cat gpurun_out/r2_1_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_operator.py -x -q -m gpu > gpurun_out/r2_1_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2_1_rest.log
tail -5 gpurun_out/r2_1_rest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_fused -c 1 -o gpurun_out/r2_1_lz4only_silesia \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/r2_1_ncu.log 2>&1
echo done
