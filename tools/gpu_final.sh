#!/bin/bash
# Final round-1 evidence run (1 GPU): smoke, parity tests, bench, ncu launch list + full capture, sweeps.
set +e
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])"; tail -3 gpurun_out/bench_final.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2>&1; cut -c1-200 gpurun_out/bench_final_ref.json
echo "== sweep"; timeout 900 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads random,silesia,zeros --flags lz4,md5,both --decode > gpurun_out/sweep_final.jsonl 2> gpurun_out/sweep_final.err; cut -c1-170 gpurun_out/sweep_final.jsonl; tail -2 gpurun_out/sweep_final.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sky_ --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -3 gpurun_out/launches_final.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 3 -c 1 -o gpurun_out/prof_final -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
