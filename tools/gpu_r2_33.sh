#!/bin/bash
# round 2, run 33: gateway-queue worker that reads the next batch while it waits for the oldest one (operators.worker_loop)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_operator.py -x -q -m gpu > $O/r2_33_operator.log 2>&1; tail -2 $O/r2_33_operator.log
timeout 200 python -m skyplane_b200.harness --gpus 1 --chunks 3072 --chunk-mib 8 --pool 32 --workload random --batch 128 --slots 3 > $O/r2_33_harness_n1_b128s3.json 2> $O/r2_33_harness.err
cat $O/r2_33_harness_n1_b128s3.json; tail -2 $O/r2_33_harness.err
timeout 200 python -m skyplane_b200.harness --gpus 1 --chunks 3072 --chunk-mib 8 --pool 32 --workload random --batch 128 --slots 4 > $O/r2_33_harness_n1_b128s4.json 2>> $O/r2_33_harness.err
cat $O/r2_33_harness_n1_b128s4.json
timeout 200 python -m skyplane_b200.harness --gpus 1 --chunks 3072 --chunk-mib 8 --pool 32 --workload silesia --batch 128 --slots 4 > $O/r2_33_harness_n1_silesia_b128s4.json 2>> $O/r2_33_harness.err
cat $O/r2_33_harness_n1_silesia_b128s4.json
echo done
