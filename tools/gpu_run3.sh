#!/bin/bash
set +e
mkdir -p gpurun_out
echo "== pcie"; timeout 300 python tools/pcie_probe.py > gpurun_out/pcie.json 2> gpurun_out/pcie.err; cat gpurun_out/pcie.json; tail -2 gpurun_out/pcie.err
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "edge_lengths or mixed or golden or empty or long_matches" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck.log
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; cat gpurun_out/bench3.json; tail -5 gpurun_out/bench3.err
echo "== size sweep"; timeout 900 python tools/sweep.py --total-mib 8192 --sizes-mib 1,4,16,64 --workloads random --flags both,lz4,md5 > gpurun_out/sweep_sizes.jsonl 2> gpurun_out/sweep_sizes.err; cat gpurun_out/sweep_sizes.jsonl; tail -3 gpurun_out/sweep_sizes.err
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 3 -c 1 -o gpurun_out/prof_r3 -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_full.log
