"""CPU suite, world_size 2 (and 3) over gloo: the row-sharded data flow of the engine -- local products, all-reduce
of Z = A^T Y and of the Gram matrices, replicated Cholesky / Jacobi factors -- reproduces the single-rank oracle."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_row_sharded_algorithm_matches_oracle(tmp_path, world):
    env = dict(os.environ, OMP_NUM_THREADS="2", MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           str(ROOT / "tests" / "_gloo_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads((tmp_path / "result.json").read_text())
    assert res["world"] == world
    assert res["sigma_rel"] < 1e-10 and res["sin_u"] < 1e-8 and res["sin_v"] < 1e-8
    assert res["orth"] < 1e-12
    assert res["same_uid"] and res["replicated_identical"]
    # the consumers' sharded data flow (csrc/rom.cu with a communicator): DMDc with the control rows under the last
    # rank's block, POD with the points split; against the single-process oracle
    assert res["dmdc_b_err"] < 1e-8 and res["dmdc_eig_err"] < 1e-8 and res["dmdc_sigma_rel"] < 1e-10, res
    assert res["pod_sin_modes"] < 1e-8 and res["pod_recon_err"] < 1e-8, res
    # the panel path on exactly rank-deficient input (rank 150 under a 160-column sketch), row-sharded
    assert res["wide_sigma_rel"] < 1e-10 and res["wide_sin_u"] < 1e-8 and res["wide_sin_v"] < 1e-8 and res["wide_orth"] < 1e-12, res


def test_bench_shard_partition_is_gpu_count_independent():
    """bench.py generates A in fixed 2^19-row blocks; any rank layout must see the same global rows."""
    sys.path.insert(0, str(ROOT))
    import torch
    import bench
    rows, n = 3 * bench.BLOCK_ROWS // 2 + 5, 4
    old = bench.BLOCK_ROWS
    try:
        bench.BLOCK_ROWS = 1024
        rows = 3 * 1024 // 2 + 5
        full, _, _ = bench.make_shard(torch, torch.device("cpu"), rows, n, 0, 1, 7)
        for world in (2, 3, 4):
            parts = [bench.make_shard(torch, torch.device("cpu"), rows, n, r, world, 7) for r in range(world)]
            cat = torch.cat([p[0] for p in parts])
            assert torch.equal(cat, full)
            assert parts[0][1] == 0 and parts[-1][2] == rows
    finally:
        bench.BLOCK_ROWS = old
