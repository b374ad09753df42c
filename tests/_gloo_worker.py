"""Worker for tests/test_sharded_gloo.py: one rank of a world_size-N gloo job that runs the engine's row-sharded
algorithm (oracle/engine_model.py) with torch.distributed all-reduces standing where NCCL stands on the GPUs."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import engine_model, ref_rsvd  # noqa: E402


def main():
    out_dir = Path(sys.argv[1])
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()

    def allreduce(x):
        t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64).copy())
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.numpy()

    rng = np.random.default_rng(2025)            # same stream on every rank => same global matrix
    m, n, k, q, p = 1501, 96, 12, 5, 10          # odd row count: ragged shards
    a = rng.standard_normal((m, n))
    omega = rng.standard_normal((n, k + p))
    per = (m + world - 1) // world
    r0, r1 = rank * per, min(m, (rank + 1) * per)
    u_loc, s, vt = engine_model.engine_rsvd(a[r0:r1], k, q, p, omega, allreduce=allreduce, global_rows=m)

    # the ncclUniqueId bootstrap used by corrla_rs_b200.ShardComm: rank 0 draws, everyone receives the same bytes
    box = [os.urandom(128) if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    uid_sum = torch.tensor([float(sum(box[0]))], dtype=torch.float64)
    uid_chk = uid_sum.clone()
    dist.all_reduce(uid_chk, op=dist.ReduceOp.SUM)
    same_uid = bool(abs(uid_chk.item() - world * uid_sum.item()) < 1e-9)

    # gather U shards on rank 0 and compare with the single-rank oracle
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, u_loc))
    if rank == 0:
        u = np.zeros((m, k))
        for g0, g1, blk in gathered:
            u[g0:g1] = blk
        u0, s0, vt0 = ref_rsvd.random_svd(a, k, q, p, omega=omega)
        res = {"sigma_rel": ref_rsvd.sigma_rel_err(s0, s), "sin_u": ref_rsvd.subspace_sine(u0, u),
               "sin_v": ref_rsvd.subspace_sine(vt0.T, vt.T), "orth": float(np.max(np.abs(u.T @ u - np.eye(k)))),
               "same_uid": same_uid, "world": world}
        (out_dir / "result.json").write_text(json.dumps(res))
    # replicated factors must be bitwise identical on all ranks (all-reduce results are)
    chk = torch.from_numpy(np.concatenate([s.ravel(), vt.ravel()]).copy())
    mx, mn = chk.clone(), chk.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    if rank == 0:
        res = json.loads((out_dir / "result.json").read_text())
        res["replicated_identical"] = bool(torch.equal(mx, mn))
        (out_dir / "result.json").write_text(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
