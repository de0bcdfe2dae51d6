#!/bin/bash
# round 2, final single-GPU evidence for profiles/ after the match.any-free prober: every GPU test, bench line, launch list,
# ncu captures, sanitizers -- most important first (the call may be cut off by the GPU budget)
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_final2_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $O/r2_final2_gpu_tests.log; tail -3 $O/r2_final2_gpu_tests.log
timeout 900 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err
echo "bench rc=$?"; tail -c 700 $O/r2_bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_final.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-queue-e2e --c3-chunks 256 > $O/r2_bench_under_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o $O/r2_lz4only_silesia_128x8MiB_final \
   python tools/sweep.py --total-mib 1024 --sizes-mib 8 --workloads silesia --flags lz4 --iters 1 > $O/r2_ncu_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o $O/r2_fused_1024x8MiB_random_final \
   python tools/sweep.py --total-mib 8192 --sizes-mib 8 --workloads random --flags both --iters 1 > $O/r2_ncu_a.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:sky_fused -s 1 -c 1 -o $O/r2_fused_config3_256x16MiB_silesia_final \
   python tools/sweep.py --total-mib 4096 --sizes-mib 16 --workloads silesia --flags both --iters 1 > $O/r2_ncu_c.log 2>&1
timeout 400 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed_compressibility or long_matches" > $O/r2_compute_sanitizer_racecheck.log 2>&1
tail -3 $O/r2_compute_sanitizer_racecheck.log
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_e2ee.py -x -q -m gpu -k "twin or mixed_compressibility or boxes_equal or sealed_frames" > $O/r2_compute_sanitizer_memcheck.log 2>&1
tail -3 $O/r2_compute_sanitizer_memcheck.log
timeout 300 compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed_compressibility" > $O/r2_compute_sanitizer_synccheck.log 2>&1
tail -3 $O/r2_compute_sanitizer_synccheck.log
timeout 200 python tools/sweep.py --total-mib 2048 --sizes-mib 8 --workloads silesia,random,zeros --flags lz4,both --iters 3 > $O/r2_final2_sweep.jsonl 2> $O/r2_final2_sweep.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_final_reference_arm.json 2>> $O/r2_bench_final.err
echo done
