#!/bin/bash
# Round 2, GPU session Q (1 GPU): phase clocks of the ring Jacobi kernel (l = 110 and l = 256).
set -u
mkdir -p gpurun_out
O=gpurun_out
CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py > $O/r02q_jacobi_dbg.txt 2>&1; echo "rc=$?"; tail -32 $O/r02q_jacobi_dbg.txt
