#!/bin/bash
# round 2, run 34: gateway-queue worker, more batches in flight
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for cfg in "128 5" "128 6" "64 8"; do set -- $cfg
timeout 120 python -m skyplane_b200.harness --gpus 1 --chunks 3072 --chunk-mib 8 --pool 32 --workload random --batch $1 --slots $2 2>> $O/r2_34_harness.err | sed "s/^{/{\"batch\": $1, \"slots\": $2, /" >> $O/r2_34_harness.jsonl
done
cut -c1-200 $O/r2_34_harness.jsonl; tail -2 $O/r2_34_harness.err
echo done
