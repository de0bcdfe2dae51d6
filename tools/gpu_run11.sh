#!/bin/bash
set +e
mkdir -p gpurun_out
cp skyplane_b200/libskychunk.so gpurun_out/libskychunk_profiled.so
echo "== ncu lz4 silesia"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o gpurun_out/prof_lz4_silesia_v5 -f python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > gpurun_out/ncu_lz4.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lz4.log
