#!/bin/bash
# round 2, run 29: parsers sleep instead of spinning, parallel plan scan, pacing slack; 12 parsers by default
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_29_parity.log 2>&1
echo "parity rc=$?" >> $O/r2_29_parity.log
tail -4 $O/r2_29_parity.log
if grep -q "rc=0" $O/r2_29_parity.log; then
timeout 100 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4 --iters 3 2> $O/r2_29_sweep_default.err | sed "s/^{/{\"build\": \"default\", /" >> $O/r2_29_sweep.jsonl
for v in wait0 wait256 wait4096; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 100 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia --flags lz4 --iters 3 \
     2> $O/r2_29_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_29_sweep.jsonl
done
timeout 200 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags both,both_nopace,lz4 --iters 2 2>> $O/r2_29_sweep_default.err | sed "s/^{/{\"build\": \"default\", /" >> $O/r2_29_sweep.jsonl
for v in pace1 pace2; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 200 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags both --iters 2 \
     2> $O/r2_29_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_29_sweep.jsonl
done
cut -c1-215 $O/r2_29_sweep.jsonl
fi
echo done
