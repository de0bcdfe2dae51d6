#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "ratio|passed|failed|Error|error" gpurun_out/pytest_gpu.log | tail -15
echo "== sweep"; timeout 600 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads random,silesia,zeros --flags lz4,md5,both,both_nopace,both_excl > gpurun_out/sweep_8mib.jsonl 2> gpurun_out/sweep_8mib.err; cat gpurun_out/sweep_8mib.jsonl; tail -3 gpurun_out/sweep_8mib.err
echo "== bench"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; cat gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 3 -c 1 -o gpurun_out/prof_r2 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
