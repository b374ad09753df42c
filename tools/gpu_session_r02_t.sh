#!/bin/bash
# Round 2, GPU session T (1 GPU): ring Jacobi after the cta-scope wait / no-spill fix; cluster of 16.
set -u
mkdir -p gpurun_out
O=gpurun_out
( echo "== default (cluster 8, 7 warps per CTA)"
  CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py 2>&1 | tail -20 | head -8
  echo "== cluster 16, 4 warps per CTA"
  CORRLA_B200_JACOBI_RING_MAXC=16 CORRLA_B200_JACOBI_RING_WPC=4 CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py 2>&1 | tail -20 | head -8
) > $O/r02t_ring.txt 2>&1
cat $O/r02t_ring.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py tests/test_gpu_rom.py -m gpu -x -q 2>&1 | tail -4
