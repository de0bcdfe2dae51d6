#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
export SKYCHUNK_TRACE=1
for cfg in "256 4" "128 4" "128 8"; do set -- $cfg; echo "== sub=$1 slots=$2"; timeout 300 python tools/e2e_probe.py --sub $1 --slots $2 --batches 12 2>&1 | tee gpurun_out/e2e_probe_$1_$2.txt | grep -E "trace|gbs"; done
unset SKYCHUNK_TRACE
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench6.json 2> gpurun_out/bench6.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench6.json')); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench6.err
