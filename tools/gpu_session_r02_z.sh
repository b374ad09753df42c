#!/bin/bash
# Round 2, GPU session Z (1 GPU): nearest-neighbour shortlist on TF32 mma.sync -- stats tests, 1M x 64 timing (TF32 and FP64 pipe).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py -m gpu -x -q 2>&1 | tail -15
timeout 300 python tools/bench_knn.py 1048576 > $O/r02z_knn_tf32.json 2> $O/r02z.err; cat $O/r02z_knn_tf32.json; tail -3 $O/r02z.err
CORRLA_B200_KNN_TF32=0 timeout 300 python tools/bench_knn.py 1048576 > $O/r02z_knn_dmma.json 2>> $O/r02z.err; cat $O/r02z_knn_dmma.json
timeout 300 python tools/bench_knn.py 262144 --exact > $O/r02z_knn_256k.json 2>> $O/r02z.err; cat $O/r02z_knn_256k.json
