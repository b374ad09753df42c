#!/bin/bash
# Round 2, GPU session M (1 GPU): all tests after the triangular/symmetric skip change, C3 bench, 512k launch list.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/r02m_pytest.log
python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-other-configs > $O/r02m_bench.json 2> $O/r02m_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('$O/r02m_bench.json')); print('C3 ms', d['ms_per_step'], 'pass', d['roofline']['avg_launch_ms'], 'nonpass', d['roofline']['non_pass_ms_per_step'], 'parity', d['parity']['ok'], d['parity']['sigma_vs_golden_rel'])"
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02m_512k_plain.json 2> $O/r02m_512k_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02m_launches_512k.csv $CMD512 > $O/r02m_ncu1.log 2>&1
python tools/launch_summary.py $O/r02m_launches_512k.csv | head -14
