#!/bin/bash
# round 2, run 32: two MD5 warps per digest CTA (16 digest CTAs instead of 32 for 1024 chunks) in the fused kernel
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for w in 1 2 4; do
SKYCHUNK_MD5_WARPS=$w timeout 100 python tools/sweep.py --total-mib 8192 --sizes-mib 8 --workloads random --flags both,md5 --iters 3 2>> $O/r2_32_sweep.err | sed "s/^{/{\"md5_warps\": $w, /" >> $O/r2_32_sweep.jsonl
SKYCHUNK_MD5_WARPS=$w timeout 100 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags both --iters 2 2>> $O/r2_32_sweep.err | sed "s/^{/{\"md5_warps\": $w, /" >> $O/r2_32_sweep.jsonl
done
cut -c1-200 $O/r2_32_sweep.jsonl
echo done
