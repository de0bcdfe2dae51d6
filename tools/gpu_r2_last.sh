#!/bin/bash
# round 2, last call: every GPU test and the smoke entry on the committed tree
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests -x -q -m gpu > $O/r2_last_gpu_tests.log 2>&1; echo "rc=$?" >> $O/r2_last_gpu_tests.log; tail -3 $O/r2_last_gpu_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/r2_last_smoke.log 2>&1; tail -2 $O/r2_last_smoke.log
echo done
