#!/bin/bash
# Round 2, GPU session ZA (1 GPU): ncu of the TF32 nearest-neighbour kernel (262144 x 64, k = 72) and the launch list of the
# whole active-subspace call.
set -u
mkdir -p gpurun_out
O=gpurun_out
CORRLA_B200_KNN_VERBOSE=1 timeout 300 python tools/bench_knn.py 1048576 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02za_launches_active.csv python tools/bench_knn.py 262144 > $O/r02za_a.log 2>&1
python tools/launch_summary.py $O/r02za_launches_active.csv | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:knn_tf32 -c 1 -o $O/r02za_knn_tf32 python tools/bench_knn.py 262144 > $O/r02za_b.log 2>&1; tail -2 $O/r02za_b.log
