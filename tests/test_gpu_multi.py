"""Multi-GPU parity (needs >= 2 GPUs; skipped on a 1-GPU box): rows sharded over ranks, NCCL all-reduce of Z and
of the Gram matrices, results against the single-process oracle."""
import json
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_rsvd_over_nccl(tmp_path, world):
    if gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           str(ROOT / "tests" / "_nccl_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads((tmp_path / "result.json").read_text())
    for name in ("gauss_rowmajor", "lowrank_colmajor", "tiny_rank_deficient", "wide_nccl_fallback", "panels_l160",
                 "panels_rank150"):
        c = res[name]
        assert c["sigma_rel"] < 1e-10, (name, c)
        assert c["sin_u"] < 1e-8 and c["sin_v"] < 1e-8, (name, c)
        assert c["orth_u"] < 1e-12, (name, c)
        assert c["sigma_tail_abs"] < 1e-9, (name, c)
        assert c["device_vs_host_sigma"] == 0.0 and c["device_vs_host_u"] == 0.0, (name, c)
    assert res["gauss_rowmajor"]["p2p_exchanges"] > 0            # sums fused into the reduction kernel over peer memory
    d = res["dmdc"]
    assert d["b_err"] < 1e-8 and d["b_vs_oracle"] < 1e-8 and d["eig_err"] < 1e-8 and d["op_err"] < 1e-8, d
    assert d["a_til_replicated"] == 0.0 and d["sigma_rel"] < 1e-10, d
    pd_ = res["pod"]
    assert pd_["sin_modes"] < 1e-8 and pd_["orth"] < 1e-12 and pd_["weights_err"] < 1e-10 and pd_["recon_err"] < 1e-8, pd_
    assert pd_["weights_replicated"] == 0.0, pd_
    cv = res["cov"]
    assert cv["cov_err"] < 1e-11 and cv["cor_err"] < 1e-11 and cv["mean_err"] < 1e-11 and cv["eig_err"] < 1e-11, cv
    assert cv["replicated"] == 0.0, cv
    st = res["streamed"]
    assert st["min_chunks"] >= 2 and st["sigma_rel"] < 1e-12 and st["u_diff"] < 1e-10 and st["vt_diff"] < 1e-10, st
    assert res["thin_q"]["orth"] < 1e-13 and res["thin_q"]["span"] < 1e-13
    rb = res["robust_stage"]
    assert rb["robust_qr_stages"] >= 1, rb                       # the sketch-preconditioned stage really ran, sharded
    assert rb["sigma_lead_rel"] < 1e-9 and rb["orth_u"] < 1e-12 and rb["orth_v"] < 1e-12, rb
    assert rb["repeat_u_diff"] == 0.0 and rb["repeat_s_diff"] == 0.0, rb     # bit-reproducible call after call


def test_peer_exchange_timeout_fails_on_every_rank(tmp_path):
    """One rank is given a 1-cycle exchange timeout and its peer arrives late: the call must fail with CORRLA_ERR_COMM
    on BOTH ranks (the poisoned epoch tells the late rank), with timings == NULL; a fresh communicator then works."""
    if gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(free_port()),
           str(ROOT / "tests" / "_nccl_timeout_worker.py"), str(tmp_path)]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-3000:])
    res = json.loads((tmp_path / "timeout.json").read_text())
    assert res["p2p"] is True
    assert res["status"] == [-6, -6], res                        # CORRLA_ERR_COMM on both ranks
    assert res["after_sigma_rel"] < 1e-10 and res["after_status"] == [0, 0], res
