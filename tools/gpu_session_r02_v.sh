#!/bin/bash
# Round 2, GPU session V (8 GPUs): the scaling run of the final build -- bench at N = 8, 4, 2, 1 the way the driver launches
# it, the reference arm under torchrun, the multi-GPU parity tests (world 2, 4, 8, peer-exchange timeout).
set -u
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > $O/r02v_topo.txt 2>&1
run() {  # n, port, extra flags...
  local n=$1 port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"
}
run 8 29701 --steps 20 --warmup 5 > $O/r02v_bench_8gpu.json 2> $O/r02v_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 600 $O/r02v_bench_8gpu.err
run 4 29702 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02v_bench_4gpu.json 2> $O/r02v_bench_4gpu.err; echo "bench4 rc=$?"; tail -c 300 $O/r02v_bench_4gpu.err
run 2 29703 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02v_bench_2gpu.json 2> $O/r02v_bench_2gpu.err; echo "bench2 rc=$?"; tail -c 300 $O/r02v_bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02v_bench_1gpu.json 2> $O/r02v_bench_1gpu.err; echo "bench1 rc=$?"
python - <<'PY'
import json
for n in (1, 2, 4, 8):
    try:
        j = json.loads(open(f"gpurun_out/r02v_bench_{n}gpu.json").read().strip().splitlines()[-1])
        e = j.get("e2e") or {}
        print(n, "ms", round(j["ms_per_step"], 3), "value", round(j["value"], 1), "e2e ms", e.get("ms_per_step"), "h2d GB/s agg", e.get("h2d_gbs_aggregate"),
              "numa", e.get("numa_binding_rank0"), "parity", (j.get("parity") or {}).get("ok_all_ranks"), "clocks", j.get("clocks", {}).get("sm_mhz"))
    except Exception as ex:
        print(n, "ERR", ex)
PY
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02v_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -6 $O/r02v_pytest_multi.log
run 8 29704 --impl reference --steps 1 --warmup 0 > $O/r02v_bench_reference_8gpu.json 2> $O/r02v_bench_ref.err; echo "ref8 rc=$?"; cut -c1-300 $O/r02v_bench_reference_8gpu.json
