#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
for v in default w24_h12 w24_h12_s4; do
  echo "== variant $v"
  if [ "$v" = "default" ]; then unset SKYCHUNK_LIB; else export SKYCHUNK_LIB=$PWD/tools/bin/libskychunk_$v.so; fi
  timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random --flags md5,both,both_excl,both_nopace > gpurun_out/sweep_v3_$v.jsonl 2> gpurun_out/sweep_v3_$v.err
  timeout 600 python tools/sweep.py --total-mib 8192 --sizes-mib 1 --workloads random,silesia --flags md5,lz4,both,both_nopace >> gpurun_out/sweep_v3_$v.jsonl 2>> gpurun_out/sweep_v3_$v.err
  python - <<PY
import json
for l in open('gpurun_out/sweep_v3_$v.jsonl'):
    d=json.loads(l); print('$v', d['workload'], d['chunk_mib'], d['flags'], 'ms %.2f'%d['kernel_ms'], 'GB/s %.1f'%d['raw_input_gbs'], 'ratio', d['ratio'] and round(d['ratio'],3))
PY
  tail -2 gpurun_out/sweep_v3_$v.err
done
