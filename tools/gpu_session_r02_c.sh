#!/bin/bash
# Round 2, GPU session C (1 GPU): super-chunk k-group path (unit tool + tests), C1 timing with phase breakdown, C5 line.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/test_gemm speed > $O/r02c_test_gemm.log 2>&1; echo "test_gemm rc=$?"; grep -c PASS $O/r02c_test_gemm.log; grep -E "FAIL|failures|SPEED C5|SPEED Gram|SPEED apply" $O/r02c_test_gemm.log
python -m pytest tests -m gpu -x -q > $O/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02c_pytest.log
python tools/time_c1.py > $O/r02c_c1.log 2>&1; cat $O/r02c_c1.log
python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak > $O/r02c_c5.json 2> $O/r02c_c5.err; echo "c5 rc=$?"; python -c "
import json; d=json.load(open('$O/r02c_c5.json')); print('C5 ms', d['ms_per_step'], 'pass avg', d['roofline']['avg_launch_ms'], 'frac', d['roofline']['frac'])"
