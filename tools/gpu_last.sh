#!/bin/bash
# last bounded check of the committed kernel: smoke + bench line + launch list
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout 120 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/smoke.log
echo "== bench"; timeout 240 python bench.py > gpurun_out/bench_last.json 2> gpurun_out/bench_last.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_last.json')); print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'], 'kernel_ms', d['roofline']['kernel_ms'], d['clocks'])"; tail -2 gpurun_out/bench_last.err
echo "== ncu launches"; timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sky_ --csv --log-file gpurun_out/launches_last.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -2 gpurun_out/launches_last.csv
