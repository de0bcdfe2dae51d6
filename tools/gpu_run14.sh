#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== decode sweep"; timeout 900 python tools/sweep.py --total-mib 2048 --sizes-mib 8,1 --workloads random,silesia,zeros --flags both --decode > gpurun_out/sweep_decode.jsonl 2> gpurun_out/sweep_decode.err; cut -c1-220 gpurun_out/sweep_decode.jsonl; tail -3 gpurun_out/sweep_decode.err
