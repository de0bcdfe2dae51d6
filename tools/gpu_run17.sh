#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== racecheck"; timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_decode.py -m gpu -x -q -k "golden or empty or mixed or long_matches or malformed or stage_flags" > gpurun_out/racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck.log
echo "== synccheck"; timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or mixed" > gpurun_out/synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/synccheck.log
echo "== size sweep (config 5, one GPU)"; timeout 900 python tools/sweep.py --total-mib 8192 --sizes-mib 1,4,16,64 --workloads random,silesia --flags both > gpurun_out/sweep_sizes_final.jsonl 2> gpurun_out/sweep_sizes.err; cut -c1-200 gpurun_out/sweep_sizes_final.jsonl; tail -2 gpurun_out/sweep_sizes.err
echo "== bench silesia 16MiB (config 3)"; timeout 1200 python bench.py --steps 3 --warmup 3 --workload silesia --chunk-mib 16 --chunks 1024 --e2e-batch 128 --cpu-chunks 512 > gpurun_out/bench_config3_final.json 2> gpurun_out/bench_config3.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config3_final.json')); print('value',d['value'],'e2e',d['e2e']['value'],'cpu',d['cpu_baseline']['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'ratio', d['config']['compression_ratio'], 'cpu ratio', d['cpu_baseline']['ratio'])"; tail -3 gpurun_out/bench_config3.err
