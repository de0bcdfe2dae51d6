#!/bin/bash
# round 2, run 26: prober without match.any (hash compares 3 / 4 / 8 lanes apart + atomic-max table updates)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_26_parity.log 2>&1
echo "parity rc=$?" >> $O/r2_26_parity.log
tail -6 $O/r2_26_parity.log
if grep -q "rc=0" $O/r2_26_parity.log; then
for v in pa10 pa11_r2 pa12_r2; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 100 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random --flags lz4 --iters 3 \
     2> $O/r2_26_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_26_sweep.jsonl
done
SKYCHUNK_LIB=tools/bin/libskychunk_pa10.so timeout 100 python tools/sweep.py --total-mib 4096 --sizes-mib 16 --workloads silesia --flags both --iters 2 \
     2>> $O/r2_26_sweep_pa10.err | sed "s/^{/{\"build\": \"pa10\", /" >> $O/r2_26_sweep.jsonl
cut -c1-260 $O/r2_26_sweep.jsonl
fi
echo done
