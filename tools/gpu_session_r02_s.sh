#!/bin/bash
# Round 2, GPU session S (1 GPU): ncu source-level capture of the ring Jacobi kernel (l = 110) and of chol_inv_kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/profile_jacobi.py > $O/r02s_plain.txt 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'jacobi_ring|chol_inv' -s 6 -c 4 -o $O/r02s_jacobi_ring python tools/profile_jacobi.py > $O/r02s_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/r02s_ncu.log
