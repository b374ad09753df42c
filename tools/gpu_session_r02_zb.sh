#!/bin/bash
# Round 2, GPU session ZB (1 GPU): exact kNN kernel with one query per warp for short query lists -- stats tests, 1M timing.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_stats.py -m gpu -x -q 2>&1 | tail -4
CORRLA_B200_KNN_VERBOSE=1 timeout 300 python tools/bench_knn.py 1048576 > $O/r02zb_knn.json 2> $O/r02zb.err; cat $O/r02zb_knn.json; grep "knn\]" $O/r02zb.err | tail -1
CORRLA_B200_KNN_VERBOSE=1 timeout 600 python tools/bench_rom.py --model active --no-cpu > $O/r02zb_active.json 2>> $O/r02zb.err; cut -c1-330 $O/r02zb_active.json
