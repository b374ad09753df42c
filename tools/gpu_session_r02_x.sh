#!/bin/bash
# Round 2, GPU session X (1 GPU): ring Jacobi hand-over by per-lane flags (A/B), the mid-iteration failure test, C1 with the
# one-pass in-loop QR, DMDc A/B of the in-loop QR.
set -u
mkdir -p gpurun_out
O=gpurun_out
( for lf in 0 1; do echo "== lane flags $lf"; CORRLA_B200_JACOBI_LANE_FLAGS=$lf CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py 2>&1 | tail -12 | head -4; done ) > $O/r02x_ring.txt 2>&1; cat $O/r02x_ring.txt
CORRLA_B200_JACOBI_LANE_FLAGS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_robustness.py -m gpu -x -q 2>&1 | tail -3
timeout 600 python -m pytest tests/test_gpu_robustness.py tests/test_gpu_fused_small.py -m gpu -x -q 2>&1 | tail -3
python tools/time_c1.py 2>&1 | tail -4
for v in 0 1; do CORRLA_B200_INLOOP_CHOLQR2=$v timeout 600 python tools/bench_rom.py --model dmdc --steps 3 --warmup 1 --no-cpu 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    j=json.loads(l); print('dmdc cholqr2=$v', round(j['ms_per_call'],1), round(j.get('pass_ms_sum',0),1), j.get('gpu_launches'))"; done
