#!/bin/bash
# Round 2, GPU session P (1 GPU): the ring Jacobi kernel (jacobi_ring.cu) and grad_at -- full GPU test suite, then the
# launch list of a 512k-row RSVD (the per-GPU shape of C3 on 8 GPUs) and the wide / c5 workloads.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02p_pytest.log
CMD512="python bench.py --rows 524288 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-peak"
timeout 300 $CMD512 > $O/r02p_512k_plain.json 2> $O/r02p_512k_plain.err; echo "512k rc=$?"
python - <<'PY'
import json
j=json.loads(open("gpurun_out/r02p_512k_plain.json").read().strip().splitlines()[-1])
print("512k ms", j["ms_per_step"], j.get("step_detail"), j.get("parity",{}).get("ok"))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02p_launches_512k.csv $CMD512 > $O/r02p_ncu.log 2>&1
python tools/launch_summary.py $O/r02p_launches_512k.csv | head -14
for wl in wide c5; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-peak > $O/r02p_${wl}.json 2>> $O/r02p_512k_plain.err
  python - <<PY
import json
j=json.loads(open("gpurun_out/r02p_${wl}.json").read().strip().splitlines()[-1])
print("${wl} ms", j["ms_per_step"], j.get("step_detail"), j.get("parity",{}).get("ok"))
PY
done
CORRLA_B200_JACOBI_RING=0 timeout 300 python bench.py --workload wide --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-peak --no-parity > $O/r02p_wide_noring.json 2>> $O/r02p_512k_plain.err
python -c "
import json
j=json.loads(open('gpurun_out/r02p_wide_noring.json').read().strip().splitlines()[-1]); print('wide (old jacobi) ms', j['ms_per_step'], j.get('step_detail'))"
