#!/bin/bash
set +e
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
echo "== bench N=$N"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench12_n$N.json 2> gpurun_out/bench12_n$N.err; echo "rc=$?"; tail -1 gpurun_out/bench12_n$N.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"; tail -4 gpurun_out/bench12_n$N.err
echo "== harness N=$N"; timeout 600 python -m skyplane_b200.harness --gpus $N --chunks $((1024*N)) --chunk-mib 8 --pool 64 --batch 128 --workload mixed > gpurun_out/harness12_n$N.json 2> gpurun_out/harness12.err; cat gpurun_out/harness12_n$N.json; tail -3 gpurun_out/harness12.err
