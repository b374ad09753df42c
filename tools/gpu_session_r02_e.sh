#!/bin/bash
# Round 2, GPU session E (1 GPU): Householder rewrite + GEMM-form kNN (tests, C1 timing, kNN timings), ncu of the triangular apply.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02e_pytest.log
python tools/time_c1.py > $O/r02e_c1.log 2>&1; cat $O/r02e_c1.log
python tools/bench_knn.py 262144 > $O/r02e_knn_256k.json 2> $O/r02e_knn.err; cat $O/r02e_knn_256k.json; tail -3 $O/r02e_knn.err
python tools/bench_knn.py 1048576 > $O/r02e_knn_1m.json 2>> $O/r02e_knn.err; cat $O/r02e_knn_1m.json
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02e_512k_plain.json 2> $O/r02e_512k_plain.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"\(int\)7, \(int\)2>" -s 2 -c 1 -o $O/r02e_apply_tri $CMD512 > $O/r02e_ncu1.log 2>&1
tail -2 $O/r02e_ncu1.log
