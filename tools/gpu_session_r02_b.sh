#!/bin/bash
# Round 2, GPU session B (1 GPU): GEMM unit tool (k-group path), tests, bench, C5 launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 ./tools/test_gemm speed > $O/r02b_test_gemm.log 2>&1; echo "test_gemm rc=$?"; tail -16 $O/r02b_test_gemm.log
python -m pytest tests -m gpu -x -q > $O/r02b_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 $O/r02b_pytest.log
python bench.py --steps 5 --warmup 3 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench rc=$?"; tail -c 600 $O/r02b_bench.err
CMDC5="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMDC5 > $O/r02b_c5_plain.json 2> $O/r02b_c5_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02b_launches_c5.csv $CMDC5 > $O/r02b_ncu2.log 2>&1
ls -la $O | tail -8
