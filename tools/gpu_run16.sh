#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (tail-lead pacing, rewritten mirrors)"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for v in default e5 e3 e3_l24 e2_l16; do
  if [ "$v" = "default" ]; then unset SKYCHUNK_LIB; else export SKYCHUNK_LIB=$PWD/tools/bin/libskychunk_$v.so; fi
  timeout 300 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia --flags lz4,both > gpurun_out/sweep_ext_$v.jsonl 2> gpurun_out/sweep_ext.err
  python - <<PY
import json
for l in open('gpurun_out/sweep_ext_$v.jsonl'):
    d=json.loads(l); print('$v', d['workload'], d['flags'], 'ms %.2f'%d['kernel_ms'], 'GB/s %.1f'%d['raw_input_gbs'], 'ratio', d['ratio'] and round(d['ratio'],4))
PY
  tail -1 gpurun_out/sweep_ext.err
done
