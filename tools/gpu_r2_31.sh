#!/bin/bash
# round 2, run 31: digest role per SM (one CTA with 4 MD5 warps, its neighbour parked) instead of one MD5 warp in each of 32 CTAs
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_31_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $O/r2_31_gpu_tests.log
tail -4 $O/r2_31_gpu_tests.log
if grep -q "rc=0" $O/r2_31_gpu_tests.log; then
timeout 200 python tools/sweep.py --total-mib 8192 --sizes-mib 8 --workloads random --flags both,md5,lz4 --iters 3 2> $O/r2_31_sweep.err >> $O/r2_31_sweep.jsonl
timeout 200 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags both,md5,lz4 --iters 2 2>> $O/r2_31_sweep.err >> $O/r2_31_sweep.jsonl
timeout 100 python tools/sweep.py --total-mib 2048 --sizes-mib 1 --workloads silesia,random --flags both --iters 3 2>> $O/r2_31_sweep.err >> $O/r2_31_sweep.jsonl
cut -c1-215 $O/r2_31_sweep.jsonl
fi
echo done
