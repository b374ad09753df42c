#!/bin/bash
# Round 2, GPU session R (1 GPU): cost of a cross-CTA hop in the ring Jacobi kernel (l = 30: 15 warps on 1, 2, 4, 8 CTAs).
set -u
mkdir -p gpurun_out
O=gpurun_out
for w in 16 8 4 2; do
  echo "== warps per CTA <= $w"
  CORRLA_B200_JACOBI_RING_WPC=$w CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py 20 2>&1 | tail -22 | grep -v "^ok" | tail -9
done > $O/r02r_ring_hops.txt 2>&1
cat $O/r02r_ring_hops.txt
