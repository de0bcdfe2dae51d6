#!/bin/bash
# final check of the committed main: every GPU test, racecheck of the polling word fix, fresh default bench line
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -x -q -m gpu > $O/r2_23_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> $O/r2_23_gpu_tests.log; tail -3 $O/r2_23_gpu_tests.log
timeout 600 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mixed_compressibility or long_matches" > $O/r2_compute_sanitizer_racecheck.log 2>&1
tail -3 $O/r2_compute_sanitizer_racecheck.log
timeout 900 python bench.py > $O/r2_bench_final.json 2> $O/r2_bench_final.err
echo "bench rc=$?"; tail -c 400 $O/r2_bench_final.json
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_23_smoke.log 2>&1; tail -2 $O/r2_23_smoke.log
echo done
