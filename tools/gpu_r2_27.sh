#!/bin/bash
# round 2, run 27: parser-count variants of the match.any-free kernel, fused config 3, one ncu capture
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > $O/r2_27_parity.log 2>&1
echo "parity rc=$?" >> $O/r2_27_parity.log
tail -4 $O/r2_27_parity.log
if grep -q "rc=0" $O/r2_27_parity.log; then
for v in pa10 pa12_r2 pa12_r1 pa13_r1 pa14_r0; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 100 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia --flags lz4 --iters 3 \
     2> $O/r2_27_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_27_sweep.jsonl
done
for v in pa12_r2 pa13_r1; do
  SKYCHUNK_LIB=tools/bin/libskychunk_$v.so timeout 200 python tools/sweep.py --total-mib 16384 --sizes-mib 16 --workloads silesia --flags both,lz4 --iters 2 \
     2>> $O/r2_27_sweep_$v.err | sed "s/^{/{\"build\": \"$v\", /" >> $O/r2_27_sweep.jsonl
done
cut -c1-230 $O/r2_27_sweep.jsonl
SKYCHUNK_LIB=tools/bin/libskychunk_pa12_r2.so timeout 400 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o $O/r2_27_lz4only_silesia_pa12 \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > $O/r2_27_ncu.log 2>&1
tail -3 $O/r2_27_ncu.log
fi
echo done
