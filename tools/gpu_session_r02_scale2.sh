#!/bin/bash
# Round 2, final 2-GPU session: bench at N = 2 and N = 1 on the same box, multi-GPU parity tests (world 2 + exchange timeout).
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29803 bench.py --gpus 2 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02s2_bench_2gpu.json 2> $O/r02s2_bench_2gpu.err; echo "bench2 rc=$?"
timeout 240 python bench.py --gpus 1 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline --no-e2e > $O/r02s2_bench_1gpu.json 2> $O/r02s2_bench_1gpu.err; echo "bench1 rc=$?"
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/r02s2_pytest_multi.log 2>&1; echo "pytest multi rc=$?"; tail -4 $O/r02s2_pytest_multi.log
python - <<'PY'
import json
for n in (1, 2):
    try:
        j = json.loads(open(f"gpurun_out/r02s2_bench_{n}gpu.json").read().strip().splitlines()[-1]); e = j.get("e2e") or {}
        print(n, "ms", round(j["ms_per_step"], 3), "e2e", e.get("ms_per_step"), "parity", (j.get("parity") or {}).get("ok_all_ranks"))
    except Exception as ex: print(n, "ERR", ex)
PY
