#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu (host path changed)"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench4.json')); print('value',d['value'],'e2e',d['e2e'],'cpu',d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench4.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench4_ref.json 2>&1; cat gpurun_out/bench4_ref.json | cut -c1-400
for v in default w24_h12 w24_h11 w32_h11 w28_h12; do
  echo "== variant $v"
  if [ "$v" = "default" ]; then unset SKYCHUNK_LIB; else export SKYCHUNK_LIB=$PWD/tools/bin/libskychunk_$v.so; fi
  timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random --flags lz4,both,md5 > gpurun_out/sweep_var_$v.jsonl 2> gpurun_out/sweep_var_$v.err
  python - <<PY
import json
for l in open('gpurun_out/sweep_var_$v.jsonl'):
    d=json.loads(l); print('$v', d['workload'], d['flags'], 'ms %.2f'%d['kernel_ms'], 'GB/s %.1f'%d['raw_input_gbs'], 'ratio', d['ratio'] and round(d['ratio'],3))
PY
  tail -2 gpurun_out/sweep_var_$v.err
done
unset SKYCHUNK_LIB
echo "== ncu lz4 silesia"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o gpurun_out/prof_lz4_silesia -f python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/ncu_lz4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lz4.log
