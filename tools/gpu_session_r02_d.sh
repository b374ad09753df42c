#!/bin/bash
# Round 2, GPU session D (1 GPU): fused-small Jacobi rewrite (tests + C1 timing), ncu of the triangular apply at 512k rows.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests/test_gpu_fused_small.py tests/test_gpu_parity.py -m gpu -x -q > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02d_pytest.log
python tools/time_c1.py > $O/r02d_c1.log 2>&1; cat $O/r02d_c1.log
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02d_512k_plain.json 2> $O/r02d_512k_plain.err &&
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"7, 2>" -s 2 -c 1 -o $O/r02d_apply_tri $CMD512 > $O/r02d_ncu1.log 2>&1
tail -3 $O/r02d_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:chol_inv -s 2 -c 1 -o $O/r02d_chol $CMD512 > $O/r02d_ncu2.log 2>&1
tail -3 $O/r02d_ncu2.log
