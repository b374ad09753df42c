#!/bin/bash
# Round 2, GPU session ZE (1 GPU): ncu --set full of the final TF32 nearest-neighbour kernel (N = 262 144) and of the final ring
# Jacobi kernel (l = 110).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 100 python tools/bench_knn.py 262144 > $O/r02ze_plain.json 2> $O/r02ze.err; echo "plain rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:knn_tf32 -c 1 -o $O/r02ze_knn_tf32 python tools/bench_knn.py 262144 > $O/r02ze_a.log 2>&1; tail -1 $O/r02ze_a.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jacobi_ring -s 1 -c 1 -o $O/r02ze_jacobi_ring python tools/profile_jacobi.py > $O/r02ze_b.log 2>&1; tail -1 $O/r02ze_b.log
