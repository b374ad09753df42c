#!/bin/bash
# Round 2: last check of the final tree on one GPU -- the whole GPU test suite and smoke().
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02fc_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02fc_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02fc_smoke.log 2>&1; echo "smoke rc=$?"; cat $O/r02fc_smoke.log
