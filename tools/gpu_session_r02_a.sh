#!/bin/bash
# Round 2, GPU session A (1 GPU): tests, bench with golden sigma, launch list at 512k rows, ncu of the C5 kernels.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/r02a_pytest.log
tail -5 $O/r02a_pytest.log
python bench.py --steps 5 --warmup 3 --write-golden $O/bench_c3_sigma.npz > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench rc=$?"
tail -c 1500 $O/r02a_bench.err
python bench.py --impl reference --steps 1 --warmup 0 > $O/r02a_bench_ref.json 2>> $O/r02a_bench.err
# launch list of one call on a 512k-row shard (what every rank of an 8-GPU run executes, minus the exchanges)
CMD512="python bench.py --rows 524288 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMD512 > $O/r02a_512k_plain.json 2> $O/r02a_512k_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_512k.csv $CMD512 > $O/r02a_ncu1.log 2>&1
# C5: launch list + full capture of its GEMM kernels
CMDC5="python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity --no-other-configs --no-peak"
$CMDC5 > $O/r02a_c5_plain.json 2> $O/r02a_c5_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02a_launches_c5.csv $CMDC5 > $O/r02a_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:skinny_gemm -s 40 -c 12 -o $O/r02a_c5_gemm $CMDC5 > $O/r02a_ncu3.log 2>&1
ls -la $O | tail -20
