#!/bin/bash
# Round 2, GPU session G (1 GPU): fused-small CholeskyQR2 (tests + C1 timing), ncu of the GEMM-form kNN kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_fused_small.py tests/test_gpu_parity.py tests/test_gpu_stats.py -m gpu -x -q > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02g_pytest.log
python tools/time_c1.py > $O/r02g_c1.log 2>&1; cat $O/r02g_c1.log
CMDK="python tools/bench_knn.py 65536"
$CMDK > $O/r02g_knn_64k.json 2> $O/r02g_knn.err && cat $O/r02g_knn_64k.json &&
ncu --set full --clock-control none --import-source on -k regex:knn_gemm -c 1 -o $O/r02g_knn_gemm $CMDK > $O/r02g_ncu1.log 2>&1
tail -2 $O/r02g_ncu1.log
