#!/bin/bash
# Round 2, final scaling session on 8 GPUs (trimmed to the GPU-minutes left): bench at N = 8 and N = 4 the way the driver
# launches it; every run carries the in-run parity block on every rank.
set -u
mkdir -p gpurun_out
O=gpurun_out
run() { local n=$1 port=$2; shift 2
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@"; }
run 8 29801 --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02s8_bench_8gpu.json 2> $O/r02s8_bench_8gpu.err; echo "bench8 rc=$?"
run 4 29802 --steps 10 --warmup 5 --no-other-configs --no-cpu-baseline > $O/r02s8_bench_4gpu.json 2> $O/r02s8_bench_4gpu.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for n in (4, 8):
    try:
        j = json.loads(open(f"gpurun_out/r02s8_bench_{n}gpu.json").read().strip().splitlines()[-1]); e = j.get("e2e") or {}
        print(n, "ms", round(j["ms_per_step"], 3), "e2e", e.get("ms_per_step"), e.get("breakdown_ms_rank0"), "parity", (j.get("parity") or {}).get("ok_all_ranks"))
    except Exception as ex: print(n, "ERR", ex)
PY
