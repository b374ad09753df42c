#!/bin/bash
# Round 2, GPU session K (1 GPU): kNN with pipelined fragment loads, ROM tests, ncu of the big triangular apply.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_rom.py tests/test_gpu_stats.py -m gpu -x -q > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02k_pytest.log
python tools/bench_knn.py 1048576 > $O/r02k_knn_1m.json 2> $O/r02k_knn.err; cat $O/r02k_knn_1m.json; tail -2 $O/r02k_knn.err
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02k_512k_plain.json 2> $O/r02k_512k_plain.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"\(int\)7, \(int\)2>" -s 0 -c 1 -o $O/r02k_apply_tri $CMD512 > $O/r02k_ncu1.log 2>&1
tail -2 $O/r02k_ncu1.log
