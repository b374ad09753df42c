#!/bin/bash
# Round 2, GPU session ZC (1 GPU): ncu source-level capture of the TF32 nearest-neighbour kernel at full C5 size.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_tf32 -c 1 -o $O/r02zc_knn_tf32_1m python tools/bench_knn.py 1048576 > $O/r02zc.log 2>&1; tail -2 $O/r02zc.log
