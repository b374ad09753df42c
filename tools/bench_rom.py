#!/usr/bin/env python
"""Timing of the two reduced-order models built on the RSVD kernels at BASELINE scale, next to
the CPU oracle.  Not the driver's bench (that is bench.py, config C3); this prints one JSON line per model for
profiles/.

  dmdc : config C4 -- x is 2 097 152 x 2048 (column-major, like faer), u is 1 x 2048, p(x,t) generator of
         dmd_rom.rs:245-267 plus 1e-3 noise, n_modes 64, n_iters 4 (two RSVDs with l = 76 on the shifted views + the
         operator / mode products).
  pod  : x is 1024 x 4 194 304 (one snapshot per row: the fat C3 shape of pod_rom.rs:56), n_modes 100, the RSVD uses
         q = 10, p = 10 as PodI hard-codes.

usage: python tools/bench_rom.py [--model dmdc|pod|both] [--steps 3] [--warmup 1] [--scale 1.0] [--no-cpu]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def gen_dmd_device(torch, n_x, nt, seed):
    """Column-major n_x x nt snapshot matrix on the device, built in column blocks."""
    dev = torch.device("cuda")
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    xs = torch.arange(n_x, dtype=torch.float64, device=dev) * (10.0 / n_x)        # mat_linspace: i * delta
    ts = torch.arange(nt, dtype=torch.float64, device=dev) * (10.0 / nt)
    u = torch.exp(0.2 * ts)
    store = torch.empty((nt, n_x), dtype=torch.float64, device=dev)                 # row t = snapshot t  => x = store.T
    for t0 in range(0, nt, 64):
        t1 = min(nt, t0 + 64)
        blk = torch.sin(xs[None, :] + 0.2 * ts[t0:t1, None]) * u[t0:t1, None]
        blk += 1e-3 * torch.randn(blk.shape, dtype=torch.float64, device=dev, generator=g)
        store[t0:t1] = blk
    return store.t(), u.reshape(1, nt).contiguous()


def time_device(torch, fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def bench_dmdc(args):
    import torch
    import corrla_rs_b200 as cb
    from oracle import ref_rom
    n_x, nt, r, q = int(2_097_152 * args.scale), 2048, 64, 4
    x, u = gen_dmd_device(torch, n_x, nt, 7)
    ctx = cb.Context(0)
    ms, ops = time_device(torch, lambda: cb.dmdc_operators(x, u, r, q, seed=8, ctx=ctx), args.steps, args.warmup)
    tm = cb.last_timings()
    l = r + 12
    flops = 2 * (2 + 2 * q) * 2.0 * n_x * (nt - 1) * l + 2 * 2.0 * n_x * (nt - 1) * r      # u row neglected
    ev = np.linalg.eigvals(ops["a_til"].cpu().numpy())
    ctx.close()
    line = {"model": "dmdc", "config": f"C4: x {n_x}x{nt} column-major f64 on the device, n_u=1, n_modes={r}, n_iters={q}, "
                                       "n_oversamples=12 (l=76); 2 RSVDs + operator/mode products",
            "ms_per_call": ms, "gflops": flops / (ms * 1e-3) * 1e-9, "flops_counted": flops,
            "passes_over_snapshots": tm["passes_over_a"], "gpu_launches": tm["gpu_launches"],
            "pass_ms_sum": tm["pass_ms"], "pass_launches": tm["pass_launches"],
            "lambda_max_abs": float(np.max(np.abs(ev))), "s_til_head": ops["s_til"].cpu().numpy().ravel()[:4].tolist()}
    if not args.no_cpu:
        rows = max(4096, n_x // 16)
        xh, uh = x[:rows].cpu().numpy(), u.cpu().numpy()
        t0 = time.perf_counter()
        ref = ref_rom.DMDc(xh, uh, 1.0, r, q, rng=np.random.default_rng(8))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"kind": "port", "sample": f"first {rows} of {n_x} state rows, same n_modes/n_iters, numpy+OpenBLAS "
                                f"restatement of dmd_rom.rs; {dt * 1e3:.0f} ms measured, x{n_x / rows:.0f} linear extrapolation = "
                                f"{dt * 1e3 * n_x / rows:.0f} ms", "ms_sample": dt * 1e3, "lambda_max_abs": float(np.max(np.abs(ref.lambdas)))}
    print(json.dumps(line), flush=True)


def bench_pod(args):
    import torch
    import corrla_rs_b200 as cb
    from oracle import ref_rom
    n_snap, n_points, r = 1024, int(4_194_304 * args.scale), 100
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    x = torch.empty((n_snap, n_points), dtype=torch.float64, device="cuda")
    for c0 in range(0, n_points, 1 << 19):
        c1 = min(n_points, c0 + (1 << 19))
        x[:, c0:c1] = torch.randn((n_snap, c1 - c0), dtype=torch.float64, device="cuda", generator=g)
    ctx = cb.Context(0)
    ms, out = time_device(torch, lambda: cb.pod_modes_weights(x, r, seed=6, ctx=ctx), args.steps, args.warmup)
    tm = cb.last_timings()
    l = r + 10
    flops = (2 + 2 * 10) * 2.0 * n_points * n_snap * l + 2.0 * n_points * n_snap * r
    modes, weights, s = out
    ctx.close()
    line = {"model": "pod", "config": f"x {n_snap}x{n_points} row-major f64 Gaussian on the device (fat: RSVD on the transposed view), "
                                      f"n_modes={r}, q=10, p=10 (l=110); RSVD + weights = x * modes",
            "ms_per_call": ms, "gflops": flops / (ms * 1e-3) * 1e-9, "flops_counted": flops,
            "passes_over_snapshots": tm["passes_over_a"], "gpu_launches": tm["gpu_launches"],
            "pass_ms_sum": tm["pass_ms"], "pass_launches": tm["pass_launches"],
            "orth_err": float((modes.T @ modes - torch.eye(r, dtype=torch.float64, device="cuda")).abs().max()),
            "s_head": s.cpu().numpy().ravel()[:3].tolist()}
    if not args.no_cpu:
        cols = max(4096, n_points // 16)
        xh = x[:, :cols].cpu().numpy()
        t0 = time.perf_counter()
        ref_rom.PodI(xh, np.arange(n_snap, dtype=np.float64).reshape(-1, 1), r, rng=np.random.default_rng(6))
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"kind": "port", "sample": f"first {cols} of {n_points} points, numpy+OpenBLAS restatement of pod_rom.rs "
                                f"(incl. the RBF weight interpolants); {dt * 1e3:.0f} ms measured, x{n_points / cols:.0f} linear "
                                f"extrapolation = {dt * 1e3 * n_points / cols:.0f} ms", "ms_sample": dt * 1e3}
    print(json.dumps(line), flush=True)


def bench_active(args):
    """config C5: active-subspace sensitivity on 1 048 576 x 64 samples, linear local fits through 72 neighbours,
    n_comps = 8; f(x) = 0.5 x^T H x with a rank-8 H plus 1e-2 noise."""
    import torch
    import corrla_rs_b200 as cb
    from oracle import ref_stats
    n, d, k, nc = int(1_048_576 * args.scale), 64, 72, 8
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    x = torch.randn((n, d), dtype=torch.float64, device="cuda", generator=g)
    hb, _ = torch.linalg.qr(torch.randn((d, 8), dtype=torch.float64, device="cuda", generator=g))
    hvals = torch.linspace(4.0, 0.5, 8, dtype=torch.float64, device="cuda")
    y = 0.5 * (((x @ hb) ** 2) * hvals).sum(dim=1) + 1e-2 * torch.randn(n, dtype=torch.float64, device="cuda", generator=g)
    ctx = cb.Context(0)
    cb.active_ss_fit(x[:8192], y[:8192], 1, k, nc, ctx=ctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    fit = cb.active_ss_fit(x, y, 1, k, nc, ctx=ctx)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    ctx.close()
    comps = fit.components()
    hbn = hb.cpu().numpy()
    resid = np.linalg.norm(comps - hbn @ (hbn.T @ comps), 2)
    line = {"model": "active_ss", "config": f"C5: {n} x {d} Gaussian samples on the device, y = 0.5 x^T H x (rank-8 H) + 1e-2 noise, "
                                            f"order 1, n_nbr {k}, n_comps {nc}; kNN + batched least squares + Gram + Jacobi EVD",
            "ms_per_call": ms, "knn_pair_dims_per_s": float(n) * n * d / (ms * 1e-3),
            "sin_angle_to_true_active_subspace": float(resid), "eigenvalues_head": np.diag(fit.singular_vals_)[:10].tolist(),
            "n_deficient": fit.n_deficient}
    if not args.no_cpu:
        xh, yh = x.cpu().numpy(), y.cpu().numpy()
        est = ref_stats.PolyGradientEstimator(xh, yh, 1, k)
        m = 24
        t0 = time.perf_counter()
        for i in range(m):
            est.grad_at(xh[i * 1000])
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"kind": "port", "sample": f"{m} of {n} gradient estimates with the numpy restatement (brute-force neighbours "
                                f"+ pinv), {dt * 1e3:.0f} ms measured, x{n / m:.0f} linear extrapolation = {dt * n / m:.0f} s",
                                "ms_sample": dt * 1e3}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="both", choices=["dmdc", "pod", "both", "active"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the long dimension (smoke runs)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.model in ("dmdc", "both"):
        bench_dmdc(args)
        import torch
        torch.cuda.empty_cache()
    if args.model in ("pod", "both"):
        bench_pod(args)
    if args.model == "active":
        bench_active(args)


if __name__ == "__main__":
    main()
