#!/bin/bash
# Round 2, GPU session H (1 GPU): all tests, C1 timing (left-looking Cholesky), kNN with autonomous warps, DMDc panels.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02h_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02h_pytest.log
python tools/time_c1.py > $O/r02h_c1.log 2>&1; cat $O/r02h_c1.log
python tools/bench_knn.py 1048576 > $O/r02h_knn_1m.json 2> $O/r02h_knn.err; cat $O/r02h_knn_1m.json; tail -3 $O/r02h_knn.err
