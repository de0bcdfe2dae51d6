#!/bin/bash
# First GPU session: parity tests, MD5 chain microbenchmark, kernel sweeps, bench line, ncu captures.
set +e
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== md5 chain"; timeout 120 tools/bin/md5_chain_bench > gpurun_out/md5_chain.txt 2>&1; cat gpurun_out/md5_chain.txt
echo "== sweep"; timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads random,silesia,zeros --flags lz4,md5,md5_excl,both,both_excl > gpurun_out/sweep_8mib.jsonl 2> gpurun_out/sweep_8mib.err; cat gpurun_out/sweep_8mib.jsonl; tail -3 gpurun_out/sweep_8mib.err
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; cat gpurun_out/bench1.json; tail -5 gpurun_out/bench1.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sky_fused --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1; tail -8 gpurun_out/launches_r1.csv
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 3 -c 1 -o gpurun_out/prof_r1 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
