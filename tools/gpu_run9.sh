#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
echo "== sweep 8MiB"; timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads random,silesia,zeros --flags lz4,md5,both > gpurun_out/sweep_8mib_v5.jsonl 2> gpurun_out/sweep.err; cat gpurun_out/sweep_8mib_v5.jsonl | cut -c1-200; tail -2 gpurun_out/sweep.err
echo "== bench random"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench9.json')); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'], 'kernel_ms', d['roofline']['kernel_ms'])"; tail -3 gpurun_out/bench9.err
echo "== bench silesia 16MiB (config 3)"; timeout 1200 python bench.py --steps 3 --warmup 3 --workload silesia --chunk-mib 16 --chunks 1024 --e2e-batch 128 --cpu-chunks 512 > gpurun_out/bench9_silesia.json 2> gpurun_out/bench9_silesia.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench9_silesia.json')); print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline'], 'kernel_ms', d['roofline']['kernel_ms'], 'ratio', d['config']['compression_ratio'])"; tail -3 gpurun_out/bench9_silesia.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sky_fused --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/launches_v5.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 3 -c 1 -o gpurun_out/prof_v5 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
