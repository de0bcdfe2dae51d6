#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "ratio|passed|failed|rror" gpurun_out/pytest_gpu.log | tail -8
for v in default w24_h12 w28_h12; do
  echo "== variant $v"
  if [ "$v" = "default" ]; then unset SKYCHUNK_LIB; else export SKYCHUNK_LIB=$PWD/tools/bin/libskychunk_$v.so; fi
  timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both > gpurun_out/sweep_v2_$v.jsonl 2> gpurun_out/sweep_v2_$v.err
  python - <<PY
import json
for l in open('gpurun_out/sweep_v2_$v.jsonl'):
    d=json.loads(l); print('$v', d['workload'], d['flags'], 'ms %.2f'%d['kernel_ms'], 'GB/s %.1f'%d['raw_input_gbs'], 'ratio', d['ratio'] and round(d['ratio'],3))
PY
  tail -2 gpurun_out/sweep_v2_$v.err
done
unset SKYCHUNK_LIB
echo "== ncu lz4 silesia"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o gpurun_out/prof_lz4_silesia_v2 -f python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/ncu_lz4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lz4.log
