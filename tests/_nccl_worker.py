"""Worker for tests/test_gpu_multi.py: one rank (one GPU) of a row-sharded rsvd over NCCL, checked against the oracle."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import corrla_rs_b200 as cb  # noqa: E402
from oracle import ref_rsvd  # noqa: E402


def main():
    out_dir = Path(sys.argv[1])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    comm = cb.ShardComm(device=local)
    results = {}
    os.environ["CORRLA_B200_STREAM_ROWS"] = "0"       # the bitwise host-vs-device checks below need the unstreamed host path
    cases = {"gauss_rowmajor": (20011, 512, 30, 5, 10, "C"), "lowrank_colmajor": (16384, 300, 20, 4, 10, "F"),
             "tiny_rank_deficient": (64, 16, 8, 12, 8, "C"),
             # Z = n16 x ld = 16000 x 36 doubles exceeds the 4 MiB peer-memory half: this case takes the NCCL fallback
             "wide_nccl_fallback": (18000, 16000, 20, 2, 8, "C"),
             # l = 160 > 128: the column-panel path (csrc/wide.cuh), cross-panel projections summed over the ranks
             "panels_l160": (12000, 400, 150, 3, 10, "C"),
             # the panel path on EXACTLY rank-deficient input (rank 150 under a 160-column sketch): columns of the second
             # panel collapse in the projection; the panel is projected and factored again (identical decision on every rank)
             "panels_rank150": (9000, 300, 150, 4, 10, "C")}
    for name, (m, n, k, q, p, order) in cases.items():
        rng = np.random.default_rng(77)
        if name == "panels_rank150":
            a = rng.standard_normal((m, 150)) @ rng.standard_normal((150, n))
        elif name.startswith("panels"):
            u0, _ = np.linalg.qr(rng.standard_normal((m, n)))
            v0, _ = np.linalg.qr(rng.standard_normal((n, n)))
            a = (u0 * (10.0 * 0.99 ** np.arange(n))) @ v0.T
        elif name.startswith("lowrank"):
            u0, _ = np.linalg.qr(rng.standard_normal((m, 40)))
            v0, _ = np.linalg.qr(rng.standard_normal((n, 40)))
            a = (u0 * (10.0 * 0.95 ** np.arange(40))) @ v0.T + 1e-2 * rng.standard_normal((m, n))
        elif name.startswith("tiny"):
            a = rng.standard_normal((m, 5)) @ rng.standard_normal((5, n))          # rank 5 < l = 16
        else:
            a = rng.standard_normal((m, n))
        l = min(k + p, n)
        omega = rng.standard_normal((n, l))
        per = (m + world - 1) // world
        r0, r1 = rank * per, min(m, (rank + 1) * per)
        shard = np.asfortranarray(a[r0:r1]) if order == "F" else np.ascontiguousarray(a[r0:r1])
        # host path and device path
        u_loc, s, vt = cb.rsvd(shard, k, q, p, omega=omega, comm=comm, global_rows=m, seed=9)
        p2p_used = cb.last_timings()["p2p_exchanges"]
        ud, sd, vd = cb.rsvd(torch.from_numpy(shard).cuda(), k, q, p, omega=torch.from_numpy(omega).cuda(), comm=comm, seed=9)
        torch.cuda.synchronize()
        gathered = [None] * world
        dist.all_gather_object(gathered, (r0, r1, np.asarray(u_loc), ud.cpu().numpy()))
        if rank == 0:
            u = np.zeros((m, k)); u2 = np.zeros((m, k))
            for g0, g1, blk, blk2 in gathered:
                u[g0:g1] = blk; u2[g0:g1] = blk2
            uo, so, vo = ref_rsvd.random_svd(a, k, q, p, omega=omega)
            nz = so.ravel() > 1e-9 * so.ravel()[0]
            kk = int(nz.sum())
            results[name] = {
                "sigma_rel": float(np.max(np.abs(so.ravel()[:kk] - s.ravel()[:kk]) / so.ravel()[:kk])),
                "sigma_tail_abs": float(np.max(np.abs(s.ravel()[kk:])) if kk < k else 0.0),
                "sin_u": ref_rsvd.subspace_sine(uo[:, :kk], u[:, :kk]), "sin_v": ref_rsvd.subspace_sine(vo[:kk].T, vt[:kk].T),
                "orth_u": float(np.max(np.abs(u.T @ u - np.eye(k)))),
                "device_vs_host_sigma": float(np.max(np.abs(sd.cpu().numpy() - s))),
                "device_vs_host_u": float(np.max(np.abs(u2 - u))), "k_checked": kk, "p2p_exchanges": p2p_used}
    # fast-decaying spectrum: cond(Y) ~ 1e16 at the reference's first QR, so every QR leaves plain CholeskyQR2 for the
    # sketch-preconditioned stage, whose second pass is skipped ON THE DEVICE (a conditional exchange: the epoch
    # handshake must still happen).  Repeated calls on one communicator stress the half-buffer reuse.
    rng = np.random.default_rng(81)
    m, n, k, q, p = 24000, 192, 24, 6, 8
    uu0, _ = np.linalg.qr(rng.standard_normal((m, n)))
    vv0, _ = np.linalg.qr(rng.standard_normal((n, n)))
    sig_true = 2.0 ** -np.arange(n)
    a = (uu0 * sig_true) @ vv0.T
    omega = rng.standard_normal((n, k + p))
    per = (m + world - 1) // world
    r0, r1 = rank * per, min(m, (rank + 1) * per)
    shard = np.ascontiguousarray(a[r0:r1])
    runs = []
    for rep in range(3):
        u_loc, s, vt = cb.rsvd(shard, k, q, p, omega=omega, comm=comm, global_rows=m, seed=9)
        runs.append((np.asarray(u_loc).copy(), np.asarray(s).copy(), np.asarray(vt).copy(), cb.last_timings()["qr_third_passes"]))
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, runs[0][0], [float(np.max(np.abs(r[0] - runs[0][0]))) for r in runs],
                                      [float(np.max(np.abs(r[1] - runs[0][1]))) for r in runs]))
    if rank == 0:
        u = np.zeros((m, k))
        for g0, g1, blk, _, _ in gathered:
            u[g0:g1] = blk
        results["robust_stage"] = {
            "robust_qr_stages": int(runs[0][3]),
            "sigma_lead_rel": float(np.max(np.abs(runs[0][1].ravel()[:8] - sig_true[:8]) / sig_true[:8])),
            "orth_u": float(np.max(np.abs(u.T @ u - np.eye(k)))),
            "orth_v": float(np.max(np.abs(runs[0][2] @ runs[0][2].T - np.eye(k)))),
            "repeat_u_diff": float(max(max(g[3]) for g in gathered)), "repeat_s_diff": float(max(max(g[4]) for g in gathered))}
    # streamed host input on every rank (first product behind the copies): same factors up to summation order
    rng = np.random.default_rng(78)
    m, n, k, q, p = 20011, 256, 24, 4, 8
    a = rng.standard_normal((m, n))
    omega = rng.standard_normal((n, k + p))
    per = (m + world - 1) // world
    r0, r1 = rank * per, min(m, (rank + 1) * per)
    base = cb.rsvd(a[r0:r1].copy(), k, q, p, omega=omega, comm=comm, global_rows=m, seed=9)
    os.environ["CORRLA_B200_STREAM_ROWS"] = "1024"
    strm = cb.rsvd(a[r0:r1].copy(), k, q, p, omega=omega, comm=comm, global_rows=m, seed=9)
    chunks = cb.last_timings()["streamed_chunks"]
    os.environ["CORRLA_B200_STREAM_ROWS"] = "0"
    gathered = [None] * world
    dist.all_gather_object(gathered, (float(np.max(np.abs(base[0] - strm[0]))), chunks))
    if rank == 0:
        results["streamed"] = {"u_diff": max(g[0] for g in gathered), "min_chunks": min(g[1] for g in gathered),
                               "sigma_rel": float(np.max(np.abs(base[1] - strm[1]) / base[1])),
                               "vt_diff": float(np.max(np.abs(base[2] - strm[2])))}
    # DMDc with the state rows sharded, POD with the points sharded: against the single-process oracle
    from oracle import ref_rom
    rng = np.random.default_rng(79)
    n_x, n_u, nt, r_true = 6001, 2, 70, 5
    qq, _ = np.linalg.qr(rng.standard_normal((n_x, r_true)))
    lam = np.array([0.95, 0.9, -0.8, 0.7, 0.5])
    amat, bmat = (qq * lam) @ qq.T, qq @ rng.standard_normal((r_true, n_u))
    uu = rng.standard_normal((n_u, nt))
    xx = np.zeros((n_x, nt)); xx[:, 0] = qq @ rng.standard_normal(r_true)
    for t in range(nt - 1):
        xx[:, t + 1] = amat @ xx[:, t] + bmat @ uu[:, t]
    r = r_true + n_u
    omegas = (rng.standard_normal((nt - 1, r + 12)), rng.standard_normal((nt - 1, r + 12)))
    per = (n_x + world - 1) // world
    r0, r1 = rank * per, min(n_x, (rank + 1) * per)
    ops = cb.dmdc_operators(xx[r0:r1].copy(), uu, r, 5, omegas=omegas, comm=comm, global_rows=n_x)
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, np.asarray(ops["b"]), np.asarray(ops["u_hat"]), np.asarray(ops["a_til"])))
    if rank == 0:
        ref = ref_rom.DMDc(xx, uu, 1.0, r, 5, omegas=omegas)
        bfull, uh = np.zeros((n_x, n_u)), np.zeros((n_x, r))
        for g0, g1, bb, hh, _ in gathered:
            bfull[g0:g1] = bb; uh[g0:g1] = hh
        a0 = gathered[0][4]
        results["dmdc"] = {"b_err": float(np.max(np.abs(bfull - bmat))), "b_vs_oracle": float(np.max(np.abs(bfull - ref.b))),
                           "eig_err": float(np.max(np.abs(np.sort(np.linalg.eigvals(a0).real) - np.sort(np.concatenate([lam, np.zeros(n_u)]))))),
                           "op_err": float(np.max(np.abs(uh @ a0 @ uh.T - ref.u_hat @ ref.a_til @ ref.u_hat.T))),
                           "a_til_replicated": float(max(np.max(np.abs(g[4] - a0)) for g in gathered)),
                           "sigma_rel": float(np.max(np.abs(np.asarray(ops["s_til"]) - ref.s_til) / ref.s_til))}
    n_snap, n_points, r = 40, 9000, 6
    base = rng.standard_normal((n_snap, 8)) * (5.0 * 0.6 ** np.arange(8))
    xp = base @ np.linalg.qr(rng.standard_normal((n_points, 8)))[0].T + 1e-6 * rng.standard_normal((n_snap, n_points))
    omega = rng.standard_normal((n_snap, r + 10))
    per = (n_points + world - 1) // world
    c0, c1 = rank * per, min(n_points, (rank + 1) * per)
    modes, weights, sv = cb.pod_modes_weights(np.ascontiguousarray(xp[:, c0:c1]), r, omega=omega, comm=comm, global_rows=n_points)
    gathered = [None] * world
    dist.all_gather_object(gathered, (c0, c1, np.asarray(modes), np.asarray(weights)))
    if rank == 0:
        ref = ref_rom.PodI(xp, np.arange(n_snap, dtype=np.float64).reshape(-1, 1), r, omega=omega)
        mfull = np.zeros((n_points, r))
        for g0, g1, mm_, _ in gathered:
            mfull[g0:g1] = mm_
        w0 = gathered[0][3]
        results["pod"] = {"sin_modes": ref_rsvd.subspace_sine(ref.modes, mfull), "orth": float(np.max(np.abs(mfull.T @ mfull - np.eye(r)))),
                          "weights_err": float(np.max(np.abs(w0 - xp @ mfull))),
                          "recon_err": float(np.max(np.abs(w0 @ mfull.T - ref.mode_weights @ ref.modes.T))),
                          "weights_replicated": float(max(np.max(np.abs(g[3] - w0)) for g in gathered))}
    # covariance / correlation with the samples sharded over the ranks (means and Gram all-reduced)
    from oracle import ref_stats
    rng = np.random.default_rng(80)
    xs = rng.standard_normal((7001, 24)) @ rng.standard_normal((24, 24)) + 30.0 * rng.standard_normal((1, 24))
    per = (7001 + world - 1) // world
    r0, r1 = rank * per, min(7001, (rank + 1) * per)
    cov_l, mu_l, ev_l, vec_l = cb.cov(xs[r0:r1].copy(), "centered", evd=True, comm=comm, global_rows=7001)
    cor_l = cb.pearson_corr(xs[r0:r1].copy(), comm=comm)          # global row count by all-reduce
    gathered = [None] * world
    dist.all_gather_object(gathered, (np.asarray(cov_l), np.asarray(cor_l)))
    if rank == 0:
        c0 = ref_stats.mat_cov_centered(xs)
        results["cov"] = {"cov_err": float(np.max(np.abs(cov_l - c0)) / np.max(np.abs(c0))),
                          "cor_err": float(np.max(np.abs(cor_l - ref_stats.pearson_corr(xs)))),
                          "mean_err": float(np.max(np.abs(mu_l.ravel() - xs.mean(axis=0)))),
                          "eig_err": float(np.max(np.abs(ev_l.ravel() - np.linalg.eigvalsh(c0)[::-1])) / np.max(np.abs(c0))),
                          "replicated": float(max(np.max(np.abs(g[0] - cov_l)) + np.max(np.abs(g[1] - cor_l)) for g in gathered))}
    # thin_q sharded
    rng = np.random.default_rng(5)
    a = rng.standard_normal((9000, 48))
    per = (9000 + world - 1) // world
    r0, r1 = rank * per, min(9000, (rank + 1) * per)
    q_loc = cb.thin_q(a[r0:r1].copy(), comm=comm, global_rows=9000)
    gathered = [None] * world
    dist.all_gather_object(gathered, (r0, r1, np.asarray(q_loc)))
    if rank == 0:
        qf = np.zeros((9000, 48))
        for g0, g1, blk in gathered:
            qf[g0:g1] = blk
        results["thin_q"] = {"orth": float(np.max(np.abs(qf.T @ qf - np.eye(48)))),
                             "span": float(np.linalg.norm(a - qf @ (qf.T @ a)) / np.linalg.norm(a))}
        (out_dir / "result.json").write_text(json.dumps(results))
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
