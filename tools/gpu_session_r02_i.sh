#!/bin/bash
# Round 2, GPU session I (1 GPU): all tests (wide DMDc/POD, f32), ncu of the kNN GEMM kernel in steady state.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02i_pytest.log
CMDK="python tools/bench_knn.py 262144"
$CMDK > $O/r02i_knn_256k.json 2> $O/r02i_knn.err && cat $O/r02i_knn_256k.json &&
ncu --set full --clock-control none --import-source on -k regex:knn_gemm -c 1 -o $O/r02i_knn_gemm $CMDK > $O/r02i_ncu1.log 2>&1
tail -2 $O/r02i_ncu1.log
