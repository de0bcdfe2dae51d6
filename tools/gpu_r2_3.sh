#!/bin/bash
# round 2, GPU call: parity of the CTA-per-block compressor + speed of the tuning builds + one ncu capture
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/r2_3_parity.log 2>&1
echo "parity rc=$?" >> gpurun_out/r2_3_parity.log
tail -25 gpurun_out/r2_3_parity.log
if grep -q "rc=0" gpurun_out/r2_3_parity.log; then
for v in p10_e4096 p8_e4096 p6_e4096 p12_e3072; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both --iters 3 \
     2> gpurun_out/r2_3_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> gpurun_out/r2_3_sweep.jsonl
done
cat gpurun_out/r2_3_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_decode.py tests/test_gpu_operator.py -x -q -m gpu > gpurun_out/r2_3_rest.log 2>&1
echo "rest rc=$?" >> gpurun_out/r2_3_rest.log
tail -5 gpurun_out/r2_3_rest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sky_fused -c 1 -o gpurun_out/r2_3_lz4only_silesia \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/r2_3_ncu.log 2>&1
fi
echo done
