#!/bin/bash
# Round 2, GPU session J (1 GPU): kNN kernel experiments (ordered DMMA; without selection), POD wide test, smoke.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_rom.py tests/test_gpu_stats.py tests/test_gpu_robustness.py -m gpu -x -q > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02j_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02j_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/r02j_smoke.log
python tools/bench_knn.py 524288 > $O/r02j_knn_512k.json 2> $O/r02j_knn.err; cat $O/r02j_knn_512k.json
CORRLA_B200_KNN_DEBUG=1 python tools/bench_knn.py 524288 > $O/r02j_knn_512k_nosel.json 2>> $O/r02j_knn.err; cat $O/r02j_knn_512k_nosel.json
