#!/bin/bash
# Round 2, GPU session U (1 GPU): ring Jacobi with the default plan (cluster 16 when placeable) -- clocks, full tests, 512k bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
CORRLA_B200_JACOBI_DEBUG=1 timeout 300 python tools/profile_jacobi.py 2>&1 | tail -12 | head -5
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02u_pytest.log
CMD512="python bench.py --rows 524288 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-peak --no-parity"
timeout 300 $CMD512 > $O/r02u_512k_plain.json 2> $O/r02u_512k_plain.err; echo "512k rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02u_launches_512k.csv $CMD512 > $O/r02u_ncu.log 2>&1
python tools/launch_summary.py $O/r02u_launches_512k.csv > $O/r02u_launch_summary.txt; cat $O/r02u_launch_summary.txt
for wl in wide c5 c2; do
  timeout 300 python bench.py --workload $wl --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-peak > $O/r02u_${wl}.json 2>> $O/r02u_512k_plain.err
done
python - <<'PY'
import json
for f in ["512k_plain","wide","c5","c2"]:
    j=json.loads(open(f"gpurun_out/r02u_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", j["ms_per_step"], j.get("step_detail"), (j.get("parity") or {}).get("ok"))
PY
