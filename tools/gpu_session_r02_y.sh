#!/bin/bash
# Round 2, GPU session Y (1 GPU): chol_inv with owner-warp branches, wide path with the one-pass in-loop block QR -- full tests,
# 512k launch list, wide / c5 / c3 timings.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02y_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02y_pytest.log
F="--no-e2e --no-cpu-baseline --no-other-configs --no-peak"
CMD512="python bench.py --rows 524288 --steps 3 --warmup 2 $F --no-parity"
timeout 300 $CMD512 > $O/r02y_512k.json 2> $O/r02y.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02y_launches_512k.csv $CMD512 > $O/r02y_ncu.log 2>&1
python tools/launch_summary.py $O/r02y_launches_512k.csv > $O/r02y_launch_summary.txt; cat $O/r02y_launch_summary.txt
for wl in c3 wide c5; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 $F > $O/r02y_${wl}.json 2>> $O/r02y.err; done
python - <<'PY'
import json
for f in ["512k","c3","wide","c5"]:
    j=json.loads(open(f"gpurun_out/r02y_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(j["ms_per_step"],3), j.get("step_detail"), "parity", (j.get("parity") or {}).get("ok"))
PY
