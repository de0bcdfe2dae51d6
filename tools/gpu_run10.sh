#!/bin/bash
set +e
mkdir -p gpurun_out
N=${1:-1}
for b in 32 64 128; do
echo "== harness N=$N batch=$b random"; timeout 900 python -m skyplane_b200.harness --gpus $N --chunks $((2048*N)) --chunk-mib 8 --pool 32 --batch $b --workload random > gpurun_out/harness_n${N}_b$b.json 2> gpurun_out/harness.err; cat gpurun_out/harness_n${N}_b$b.json; tail -3 gpurun_out/harness.err
done
echo "== harness N=$N batch=64 mixed"; timeout 900 python -m skyplane_b200.harness --gpus $N --chunks $((2048*N)) --chunk-mib 8 --pool 32 --batch 64 --workload mixed > gpurun_out/harness_n${N}_mixed.json 2> gpurun_out/harness.err; cat gpurun_out/harness_n${N}_mixed.json; tail -3 gpurun_out/harness.err
