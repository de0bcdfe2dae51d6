#!/bin/bash
set +e
mkdir -p gpurun_out
N=${1:-2}
echo "== bench N=1"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench5_n1.json 2> gpurun_out/bench5_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench5_n1.json')); print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['value'])"; tail -3 gpurun_out/bench5_n1.err
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench5_n$N.json 2> gpurun_out/bench5_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench5_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'])"; tail -5 gpurun_out/bench5_n$N.err
echo "== reference arm under torchrun"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>gpurun_out/ref_n$N.err | cut -c1-300
echo "== harness N=$N"; timeout 900 python -m skyplane_b200.harness --gpus $N --chunks 1024 --chunk-mib 8 --pool 32 --batch 64 > gpurun_out/harness_n$N.json 2> gpurun_out/harness_n$N.err; cat gpurun_out/harness_n$N.json; tail -3 gpurun_out/harness_n$N.err
