#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_operator.py -x -q -m gpu > $O/r2_24_operator.log 2>&1
echo "operator tests rc=$?" >> $O/r2_24_operator.log; tail -4 $O/r2_24_operator.log
timeout 600 python tools/sweep.py --total-mib 65536 --sizes-mib 8 --workloads random --flags md5,lz4,both --iters 2 2> $O/r2_24_sweep.err > $O/r2_sweep_max_resident_8192x8MiB_random.jsonl
cut -c1-220 $O/r2_sweep_max_resident_8192x8MiB_random.jsonl; tail -2 $O/r2_24_sweep.err
echo done
