#!/bin/bash
# Round 2, GPU session L (1 GPU): kNN warp-stagger experiment.
set -u
mkdir -p gpurun_out
O=gpurun_out
for ns in 0 600 1200 2000 3000; do
  CORRLA_B200_KNN_STAGGER_NS=$ns python tools/bench_knn.py 524288 2>> $O/r02l_knn.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('stagger_ns=$ns', d['gemm_form'])"
done | tee $O/r02l_knn_stagger.txt
