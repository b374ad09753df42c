#!/bin/bash
# Round 2, GPU session ZF (1 GPU): sumsq_finalize with independent accumulators -- full suite, c3 / c5 timing, launch list line.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > $O/r02zf_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02zf_pytest.log
F="--no-e2e --no-cpu-baseline --no-other-configs --no-peak"
for wl in c3 c5; do timeout 200 python bench.py --workload $wl --steps 10 --warmup 3 $F > $O/r02zf_${wl}.json 2>> $O/r02zf.err; done
python - <<'PY'
import json
for f in ["c3","c5"]:
    j=json.loads(open(f"gpurun_out/r02zf_{f}.json").read().strip().splitlines()[-1])
    print(f, "ms", round(j["ms_per_step"],3), "parity", (j.get("parity") or {}).get("ok"), "sigma vs golden", (j.get("parity") or {}).get("sigma_vs_golden_rel"))
PY
