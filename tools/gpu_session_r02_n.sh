#!/bin/bash
# Round 2, GPU session N (1 GPU): the measurement run of the final build -- tests, smoke, bench (both arms), launch lists,
# ncu --set full of the pass kernels, the other configs, the reduced-order models.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02n_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02n_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/r02n_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r02n_bench_c3_1gpu.json 2> $O/r02n_bench.err; echo "bench rc=$?"; tail -c 400 $O/r02n_bench.err
python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 > $O/r02n_bench_reference_1gpu.json 2>> $O/r02n_bench.err; echo "ref rc=$?"
# launch list + full capture of the two pass kernels on the FULL-SIZE C3 workload
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD > $O/r02n_c3_plain.json 2> $O/r02n_c3_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02n_launches_bench_steps2_warmup1.csv $CMD > $O/r02n_ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 0 -c 2 -o $O/r02n_pass_kernels $CMD > $O/r02n_ncu_b.log 2>&1
tail -2 $O/r02n_ncu_b.log
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02n_512k_plain.json 2> $O/r02n_512k_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02n_launches_rsvd_512k_rows.csv $CMD512 > $O/r02n_ncu_c.log 2>&1
CMDC5="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMDC5 > $O/r02n_c5_plain.json 2> $O/r02n_c5_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02n_launches_c5.csv $CMDC5 > $O/r02n_ncu_d.log 2>&1
for wl in c1 c2 c5 wide; do
  python bench.py --workload $wl --steps 5 --warmup 3 --no-other-configs > $O/r02n_bench_${wl}_1gpu.json 2>> $O/r02n_bench.err; echo "$wl rc=$?"
done
python tools/bench_rom.py --model both --steps 3 --warmup 1 > $O/r02n_bench_rom.json 2> $O/r02n_bench_rom.err; echo "rom rc=$?"; tail -c 300 $O/r02n_bench_rom.err
python tools/bench_rom.py --model active > $O/r02n_bench_active_ss_c5.json 2>> $O/r02n_bench_rom.err; cat $O/r02n_bench_active_ss_c5.json
ls -la $O | grep r02n | wc -l
