#!/bin/bash
# Round 2, GPU session W (1 GPU): in-loop QR as one Cholesky pass (basis only) -- full tests, C3 / 512k / c5 / c2 / wide timings,
# A/B against CORRLA_B200_INLOOP_CHOLQR2=1, POD launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02w_pytest.log
F="--no-e2e --no-cpu-baseline --no-other-configs --no-peak"
timeout 300 python bench.py --steps 10 --warmup 3 $F > $O/r02w_c3.json 2> $O/r02w.err
CORRLA_B200_INLOOP_CHOLQR2=1 timeout 300 python bench.py --steps 10 --warmup 3 $F --no-parity > $O/r02w_c3_cholqr2.json 2>> $O/r02w.err
timeout 300 python bench.py --rows 524288 --steps 5 --warmup 3 $F > $O/r02w_512k.json 2>> $O/r02w.err
for wl in c5 c2 wide c1; do timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 $F > $O/r02w_${wl}.json 2>> $O/r02w.err; done
python - <<'PY'
import json
for f in ["c3","c3_cholqr2","512k","c5","c2","wide","c1"]:
    j=json.loads(open(f"gpurun_out/r02w_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(j["ms_per_step"],3), j.get("step_detail"), "parity", (j.get("parity") or {}).get("ok"), "launches/call", j.get("step_detail",{}).get("launches_per_call"))
PY
timeout 600 python tools/bench_rom.py --model both --steps 3 --warmup 1 > $O/r02w_bench_rom.json 2>> $O/r02w.err; cut -c1-260 $O/r02w_bench_rom.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02w_launches_pod.csv python tools/bench_rom.py --model pod --steps 1 --warmup 0 > $O/r02w_ncu_pod.log 2>&1
python tools/launch_summary.py $O/r02w_launches_pod.csv > $O/r02w_pod_summary.txt 2>&1; head -24 $O/r02w_pod_summary.txt
