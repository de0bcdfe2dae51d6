#!/bin/bash
# multi-GPU round-2 measurements on one box: host<->device copy ceilings, bench at N GPUs (both arms), gateway-queue stream
# (BASELINE config 4), chunk-size sweep on every GPU at once (config 5).  usage: gpu_multi.sh N tag [quick]
N=${1:-8}; TAG=${2:-r2_n$N}; QUICK=$3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,pci.bus_id,name --format=csv > gpurun_out/${TAG}_gpus.txt
numactl -H > gpurun_out/${TAG}_numa.txt 2>&1 || lscpu | grep -i numa > gpurun_out/${TAG}_numa.txt
ALL=$(seq -s, 0 $((N-1)))
if [ "$N" -ge 8 ]; then SETS="0 0,1,2,3 4,5,6,7 $ALL"; elif [ "$N" -ge 2 ]; then SETS="0 $ALL"; else SETS="0"; fi
timeout 300 python tools/pcie_probe.py --sets $SETS > gpurun_out/${TAG}_pcie_probe.jsonl 2> gpurun_out/${TAG}_pcie_probe.err
cut -c1-400 gpurun_out/${TAG}_pcie_probe.jsonl
STEPS=10; C3=1024; QC=1024; HC=12800
if [ -n "$QUICK" ]; then STEPS=3; C3=128; QC=128; HC=512; fi
timeout 600 python bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
tail -c 1500 gpurun_out/${TAG}_bench_reference.json
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $STEPS --warmup 3 \
    --c3-chunks $C3 --queue-chunks $QC > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/${TAG}_bench.json; tail -3 gpurun_out/${TAG}_bench.err
timeout 600 python -m skyplane_b200.harness --gpus $N --chunks $HC --chunk-mib 8 --pool 64 --workload mixed > gpurun_out/${TAG}_harness_config4.json 2> gpurun_out/${TAG}_harness.err
cat gpurun_out/${TAG}_harness_config4.json; tail -2 gpurun_out/${TAG}_harness.err
if [ -z "$QUICK" ]; then
  for g in $(seq 0 $((N-1))); do
    CUDA_VISIBLE_DEVICES=$g timeout 600 python tools/sweep.py --total-mib 8192 --sizes-mib 1,4,16,64 --workloads random,silesia --flags both,lz4 --iters 2 \
       2> gpurun_out/${TAG}_sweep5_gpu$g.err | sed "s/^{/{\"gpu\": $g, /" > gpurun_out/${TAG}_sweep5_gpu$g.jsonl &
  done
  wait
  cat gpurun_out/${TAG}_sweep5_gpu*.jsonl > gpurun_out/${TAG}_sweep_config5.jsonl
  head -8 gpurun_out/${TAG}_sweep5_gpu0.jsonl | cut -c1-200
fi
echo done
