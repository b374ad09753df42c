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
from oracle import engine_model, ref_rom, ref_rsvd  # noqa: E402


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
    # the sharded data flow of DMDc (state rows split, control rows under the last rank) and POD (points split)
    rng2 = np.random.default_rng(77)
    n_x, n_u, nt, r_true = 403, 2, 50, 4
    qq, _ = np.linalg.qr(rng2.standard_normal((n_x, r_true)))
    lam = np.array([0.9, -0.8, 0.7, 0.5])
    amat, bmat = (qq * lam) @ qq.T, qq @ rng2.standard_normal((r_true, n_u))
    uu = rng2.standard_normal((n_u, nt))
    xx = np.zeros((n_x, nt)); xx[:, 0] = qq @ rng2.standard_normal(r_true)
    for t in range(nt - 1):
        xx[:, t + 1] = amat @ xx[:, t] + bmat @ uu[:, t]
    r = r_true + n_u
    omegas = (rng2.standard_normal((nt - 1, r + 12)), rng2.standard_normal((nt - 1, r + 12)))
    per = (n_x + world - 1) // world
    d0, d1 = rank * per, min(n_x, (rank + 1) * per)
    a_til, b_loc, _ms, s_til, _uh = engine_model.dmdc_sharded(xx[d0:d1], uu, r, 5, omegas, rank, world, allreduce=allreduce,
                                                              n_x_global=n_x)
    n_snap, n_points, rp = 30, 1201, 5
    base = rng2.standard_normal((n_snap, 7)) * (4.0 * 0.5 ** np.arange(7))
    xp = base @ np.linalg.qr(rng2.standard_normal((n_points, 7)))[0].T + 1e-7 * rng2.standard_normal((n_snap, n_points))
    omega_p = rng2.standard_normal((n_snap, rp + 10))
    per = (n_points + world - 1) // world
    c0, c1 = rank * per, min(n_points, (rank + 1) * per)
    modes_loc, weights, _sp = engine_model.pod_sharded(xp[:, c0:c1], rp, omega_p, allreduce=allreduce, n_points_global=n_points)
    gathered = [None] * world
    dist.all_gather_object(gathered, (d0, d1, b_loc, c0, c1, modes_loc))
    if rank == 0:
        refd = ref_rom.DMDc(xx, uu, 1.0, r, 5, omegas=omegas)
        refp = ref_rom.PodI(xp, np.arange(n_snap, dtype=np.float64).reshape(-1, 1), rp, omega=omega_p)
        bfull, mfull = np.zeros((n_x, n_u)), np.zeros((n_points, rp))
        for g0, g1, bb, h0, h1, mm in gathered:
            bfull[g0:g1] = bb
            mfull[h0:h1] = mm
        res = json.loads((out_dir / "result.json").read_text())
        res["dmdc_b_err"] = float(np.max(np.abs(bfull - bmat)))
        res["dmdc_eig_err"] = float(np.max(np.abs(np.sort(np.linalg.eigvals(a_til).real) - np.sort(np.concatenate([lam, np.zeros(n_u)])))))
        res["dmdc_sigma_rel"] = float(np.max(np.abs(s_til - refd.s_til) / refd.s_til))
        res["pod_sin_modes"] = ref_rsvd.subspace_sine(refp.modes, mfull)
        res["pod_recon_err"] = float(np.max(np.abs(weights @ mfull.T - refp.mode_weights @ refp.modes.T)))
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
    # the panel path (l > 128) on exactly rank-deficient input, row-sharded: the cross products of the block Gram-Schmidt and
    # the panel Gram matrices are the all-reduced quantities; a panel that took the robust stage is projected again
    rng3 = np.random.default_rng(31)
    mw, nw, rw, kw, pw = 1203, 180, 150, 150, 10
    aw = rng3.standard_normal((mw, rw)) @ rng3.standard_normal((rw, nw))
    omw = rng3.standard_normal((nw, kw + pw))
    per = (mw + world - 1) // world
    w0, w1 = rank * per, min(mw, (rank + 1) * per)
    uw_loc, sw, vw = engine_model.wide_rsvd(aw[w0:w1], kw, 4, pw, omw, allreduce=allreduce, global_rows=mw)
    gathered = [None] * world
    dist.all_gather_object(gathered, (w0, w1, uw_loc))
    if rank == 0:
        uw = np.zeros((mw, kw))
        for g0, g1, blk in gathered:
            uw[g0:g1] = blk
        u0, s0, vt0 = ref_rsvd.random_svd(aw, kw, 4, pw, omega=omw)
        res = json.loads((out_dir / "result.json").read_text())
        res["wide_sigma_rel"] = ref_rsvd.sigma_rel_err(s0, sw)
        res["wide_sin_u"] = ref_rsvd.subspace_sine(u0, uw)
        res["wide_sin_v"] = ref_rsvd.subspace_sine(vt0.T, vw.T)
        res["wide_orth"] = float(np.max(np.abs(uw.T @ uw - np.eye(kw))))
        (out_dir / "result.json").write_text(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
