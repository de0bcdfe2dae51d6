#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in pa10 pa11_r2 pa10_lit32 pa10_lit8; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia --flags lz4 --iters 3 \
     2> $O/r2_19_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_19_sweep.jsonl
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 300 python tools/sweep.py --total-mib 4096 --sizes-mib 16 --workloads silesia --flags both,both_nopace --iters 2 \
     2>> $O/r2_19_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_19_sweep.jsonl
done
cut -c1-200 $O/r2_19_sweep.jsonl
echo done
