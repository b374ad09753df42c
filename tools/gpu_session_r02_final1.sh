#!/bin/bash
# Round 2, final 1-GPU measurement session of the final build: tests, smoke, bench (both arms), launch lists, the other
# configs, the reduced-order models and the active-subspace job.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02f1_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02f1_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/r02f1_smoke.log
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02f1_bench_c3_1gpu.json 2> $O/r02f1_bench.err; echo "bench rc=$?"; tail -c 300 $O/r02f1_bench.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/r02f1_bench_reference_1gpu.json 2>> $O/r02f1_bench.err; echo "ref rc=$?"
F="--no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
CMD="python bench.py --steps 2 --warmup 1 $F"
timeout 300 $CMD > $O/r02f1_c3_plain.json 2>> $O/r02f1_bench.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f1_launches_bench_steps2_warmup1.csv $CMD > $O/r02f1_ncu_a.log 2>&1
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 $F"
timeout 300 $CMD512 > $O/r02f1_512k_plain.json 2>> $O/r02f1_bench.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f1_launches_rsvd_512k_rows.csv $CMD512 > $O/r02f1_ncu_c.log 2>&1
CMDC5="python bench.py --workload c5 --steps 1 --warmup 1 $F"
timeout 300 $CMDC5 > $O/r02f1_c5_plain.json 2>> $O/r02f1_bench.err &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02f1_launches_c5.csv $CMDC5 > $O/r02f1_ncu_d.log 2>&1
for wl in c1 c2 c5 wide; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-other-configs > $O/r02f1_bench_${wl}_1gpu.json 2>> $O/r02f1_bench.err; echo "$wl rc=$?"
done
timeout 900 python tools/bench_rom.py --model both --steps 3 --warmup 1 > $O/r02f1_bench_rom.json 2> $O/r02f1_bench_rom.err; echo "rom rc=$?"
CORRLA_B200_KNN_VERBOSE=1 timeout 900 python tools/bench_rom.py --model active > $O/r02f1_bench_active_ss_c5.json 2> $O/r02f1_active.err; grep "knn\]" $O/r02f1_active.err | tail -2; cut -c1-400 $O/r02f1_bench_active_ss_c5.json
python - <<'PY'
import json
for f in ["bench_c3_1gpu","bench_c1_1gpu","bench_c2_1gpu","bench_c5_1gpu","bench_wide_1gpu","512k_plain"]:
    try:
        j=json.loads(open(f"gpurun_out/r02f1_{f}.json").read().strip().splitlines()[-1])
        print(f, "ms", round(j["ms_per_step"],3), "frac", j["roofline"].get("frac"), "e2e", (j.get("e2e") or {}).get("ms_per_step"), "parity", (j.get("parity") or {}).get("ok"))
    except Exception as ex: print(f, "ERR", ex)
for l in open("gpurun_out/r02f1_bench_rom.json").read().strip().splitlines():
    j=json.loads(l); print(j["model"], round(j["ms_per_call"],1))
PY
