#!/usr/bin/env python
"""bench.py -- RSVD wall ms & GFLOP/s (f64) on 1/2/4/8 B200, beside the CPU restatement of the reference.

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --steps K --warmup W        # CPU arm: oracle port of random_svd.rs on host cores

A "step" is one full rsvd call (2 + 2*n_iter passes over A, CholeskyQR, SVD of B, U = Q*Ub) on the BASELINE.json
workload: C3 = 4 194 304 x 1024 f64 Gaussian, n_rank 100, n_oversamples 10, n_iters 4 (l = 110), row-sharded over
the N ranks (strong scaling: total rows fixed).  `value` = (2+2q) * 2*m*n*l / wall, the metric of SURVEY section 8(d),
with A already resident in HBM; `e2e` is the same call made with HOST (pinned) buffers, copies inside the timed
region.  One JSON line on stdout (rank 0).  After the timed region the run checks itself (`parity`): orthonormality,
the residual A^T U - V S through one extra sharded pass, sigma against the committed golden vector of this seed, and a
smaller sharded case against the CPU oracle; a parity failure makes the run exit non-zero.
"""
from __future__ import annotations

import os
import sys

# The CPU arm may be launched under torch.distributed.run, which exports OMP_NUM_THREADS=1 to every rank: give the BLAS
# behind numpy all host cores back BEFORE numpy is imported (the library reads the variables when it is loaded).
if "--impl=reference" in sys.argv or ("--impl" in sys.argv and "reference" in sys.argv[sys.argv.index("--impl") + 1:sys.argv.index("--impl") + 2]):
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (rows, cols, n_rank, n_iters, n_oversamples)
    "c3": (4_194_304, 1024, 100, 4, 10),      # BASELINE.json configs[2], the north-star target
    "c2": (20_000, 20_000, 100, 4, 10),       # configs[1]
    "c5": (1_048_576, 64, 8, 8, 10),          # configs[4]
    "wide": (1_048_576, 1024, 240, 4, 16),    # not a BASELINE config: l = 256 runs as 2 column panels (csrc/wide.cuh)
    "c1": (100, 100, 10, 12, 8),              # configs[0] (README example)
}
BLOCK_ROWS = 1 << 19      # A is generated in fixed row blocks so that every GPU count sees the same matrix
TOL_SIGMA = 1e-10         # north-star tolerances: singular values, relative
TOL_ANGLE = 1e-8          # sine of the largest principal angle
GOLDEN_SIGMA = ROOT / "tests" / "golden" / "bench_c3_sigma.npz"
DMMA_ISSUE_PEAK = 37.0    # TFLOP/s: DMMA.8x8x4 issue rate at 1965 MHz (profiles/r01_fp64_pipe_microbench.txt)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override total rows (debug)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-peak", action="store_true", help="skip the in-run FP64 peak measurement")
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--write-golden", default="", help="write the sigma vector of this run to this .npz (1 GPU)")
    return ap.parse_args()


def flops_of(m, n, l, q):
    return (2 + 2 * q) * 2.0 * m * n * l


def fp64_peak_file():
    """FP64 roofline denominator from a file.  MEASURED_PEAKS.json (driver-written) holds only bf16 and HBM figures, so
    the committed FP64 number is this repo's own measurement on the same pool: torch.matmul f64 8192^3 (cuBLAS DGEMM),
    same method as the bf16 entry -- profiles/r01_cublas_fp64_marks.json.  The run also measures it live
    (measure_fp64_peak) and reports both."""
    f = ROOT / "profiles" / "r01_cublas_fp64_marks.json"
    try:
        d = json.loads(f.read_text())
        return float(d["dgemm_8192_tflops_sustained"]), "profiles/r01_cublas_fp64_marks.json (torch.matmul f64 8192^3 sustained, round-1 gpurun)"
    except Exception:
        return DMMA_ISSUE_PEAK, "fallback: DMMA.8x8x4 issue-rate peak at 1965 MHz"


def measure_fp64_peak(torch, device, seconds=1.0):
    """cuBLAS DGEMM 8192^3 through torch.matmul, back to back for about `seconds`, CUDA events: the same method
    MEASURED_PEAKS.json uses for its bf16 entry, run on THIS GPU in THIS process.  A yardstick only: no product code."""
    n = 8192
    g = torch.Generator(device=device)
    g.manual_seed(1234)
    a = torch.randn((n, n), dtype=torch.float64, device=device, generator=g)
    b = torch.randn((n, n), dtype=torch.float64, device=device, generator=g)
    c = torch.empty((n, n), dtype=torch.float64, device=device)
    for _ in range(2):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    total_ms, reps, best = 0.0, 0, 0.0
    while total_ms < seconds * 1e3 and reps < 400:
        e0.record()
        for _ in range(4):
            torch.matmul(a, b, out=c)
        e1.record()
        torch.cuda.synchronize(device)
        ms = float(e0.elapsed_time(e1))
        total_ms += ms
        reps += 4
        best = max(best, 4 * 2.0 * n ** 3 / (ms * 1e-3) * 1e-12)
    del a, b, c
    torch.cuda.empty_cache()
    return {"sustained": reps * 2.0 * n ** 3 / (total_ms * 1e-3) * 1e-12, "burst": best, "matmuls": reps,
            "seconds": total_ms * 1e-3}


def measured_traffic(rows_local, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from an `ncu --set full`
    capture committed under profiles/ (bytes per row of A, scaled to this GPU's rows); None if no capture exists."""
    f = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(f.read_text())
        if int(d["cols"]) != int(n):
            return None, None
        src = f"profiles/traffic.json: ncu --set full capture of {d.get('source', 'the pass kernel')}, " \
              f"{d.get('when', 'round 1')}; bytes per row of A scaled to this GPU's rows (not re-measured in this run)"
        return float(d["dram_bytes_per_row"]) * rows_local, src
    except Exception:
        return None, None


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of random_svd.rs on the host cores
# --------------------------------------------------------------------------------------------------
def use_all_cpu_threads():
    """Make the BLAS under numpy use every host core, whatever OMP_NUM_THREADS the launcher exported; returns the thread
    count actually in force."""
    want = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_info, threadpool_limits
        threadpool_limits(limits=want)
        n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return want


def cpu_time_oracle(sample_rows, n, k, q, p, seed, repeats=1, warmup=0):
    from oracle import ref_rsvd
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((sample_rows, n))
    omega = rng.standard_normal((n, min(k + p, n)))
    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        ref_rsvd.random_svd(a, k, q, p, omega=omega)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = use_all_cpu_threads()
    rows, n, k, q, p = WORKLOADS[args.workload]
    if args.rows:
        rows = args.rows
    l = min(k + p, n)
    # bounded sample: rows chosen so that one step is a few seconds on ~8 cores (about 70 GFLOP/s measured in the
    # survey container) and the whole K+W run stays within minutes
    budget_flops = 3.0 * 70e9
    sample = args.cpu_sample_rows or int(min(rows, max(1024, budget_flops / ((2 + 2 * q) * 2.0 * n * l))))
    sample = max(1, min(rows, (sample // 1024) * 1024 or sample))
    if sample < n:
        sample = rows              # keep the sample tall (a fat slice is a different problem)
    times = cpu_time_oracle(sample, n, k, q, p, args.seed, repeats=args.steps, warmup=args.warmup)
    ms = 1e3 * float(np.mean(times))
    value = flops_of(sample, n, l, q) / (ms * 1e-3) * 1e-9
    cfg = workload_config(args.workload, rows, n, k, q, p, args.gpus)
    # the CPU arm times a row slice of the workload, not the whole matrix: say so where the configs are compared
    cfg["cpu_sample_rows"] = sample
    cfg["cpu_sample_note"] = f"the CPU arm ran the first {sample} of {rows} rows (a rate, GFLOP/s, is reported)"
    line = {
        "impl": "reference",
        "metric": "rsvd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "host_cpus": os.cpu_count(),
                         "sample": f"first {sample} of {rows} rows x {n} cols, same k/q/p; numpy+OpenBLAS restatement "
                                   f"of random_svd.rs (oracle/ref_rsvd.py), not faer; wall {ms:.1f} ms per call; "
                                   f"linear extrapolation to {rows} rows: {ms * rows / sample:.0f} ms"},
        "e2e": {"value": value, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(name, rows, n, k, q, p, gpus):
    return {"workload": f"{name}: rsvd of {rows}x{n} f64 Gaussian, n_rank={k}, n_iters={q}, n_oversamples={p} "
                        f"(l={min(k + p, n)}), row-major, rows sharded over {gpus} GPU(s)",
            "rows": rows, "cols": n, "n_rank": k, "n_iters": q, "n_oversamples": p,
            "cache": "inputs larger than L2 (A >> 126 MB); no flush needed",
            "schedule": "reference (QR only when i > 2, Frobenius scaling each trip)"}


# --------------------------------------------------------------------------------------------------
# clocks sampler
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], 0, set(), []
        for ts, ln in self.lines:
            if ts < t0 or ts > t1 + 0.2:
                continue
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = max(smax, float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def make_shard(torch, device, rows_total, n, rank, world, seed):
    """Rows [r0, r1) of the global Gaussian matrix; block b (BLOCK_ROWS rows) is drawn from generator seed+b, so the
    matrix does not depend on the number of GPUs."""
    per = (rows_total + world - 1) // world
    r0, r1 = rank * per, min(rows_total, (rank + 1) * per)
    a = torch.empty((r1 - r0, n), dtype=torch.float64, device=device)
    b0, b1 = r0 // BLOCK_ROWS, (r1 - 1) // BLOCK_ROWS
    for b in range(b0, b1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000 + b)
        lo, hi = b * BLOCK_ROWS, min(rows_total, (b + 1) * BLOCK_ROWS)
        blk = torch.randn((hi - lo, n), dtype=torch.float64, device=device, generator=g)
        s0, s1 = max(lo, r0), min(hi, r1)
        a[s0 - r0:s1 - r0] = blk[s0 - lo:s1 - lo]
        del blk
    return a, r0, r1


class Ranks:
    """The few cross-rank helpers the bench needs (torch.distributed; no-ops on one GPU)."""

    def __init__(self, torch, dist, device, world, rank):
        self.torch, self.dist, self.device, self.world, self.rank = torch, dist, device, world, rank

    def barrier(self):
        self.torch.cuda.synchronize(self.device)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def _red(self, x, op):
        if self.world == 1:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, x):
        return self._red(x, self.dist.ReduceOp.MAX if self.world > 1 else None)

    def sum(self, x):
        return self._red(x, self.dist.ReduceOp.SUM if self.world > 1 else None)

    def sum_tensor(self, t):
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t

    def gather_rows(self, t):
        """Concatenate the ranks' row blocks (equal sizes) on every rank."""
        if self.world == 1:
            return t
        parts = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(parts, t.contiguous())
        return self.torch.cat(parts, dim=0)


def parity_block(torch, cb, rk, a, out, k, q, p, rows, n, seed, ctx, comm, workload, write_golden):
    """Self-check of the run, AFTER the timed region.  Nothing here is timed and nothing here is product code: torch
    matmuls (cuBLAS) and the CPU oracle are used as independent checkers of the engine's output."""
    u, s, vt = out
    res = {"tolerance": {"sigma_rel": TOL_SIGMA, "subspace_sine": TOL_ANGLE}}
    eye = torch.eye(k, dtype=torch.float64, device=a.device)
    gram_u = rk.sum_tensor(u.T @ u)
    res["ortho_u_max"] = float((gram_u - eye).abs().max().item())
    res["ortho_v_max"] = float((vt @ vt.T - eye).abs().max().item())
    # one extra sharded pass over A: A^T U (summed over the ranks) must equal V * diag(S)
    atu = rk.sum_tensor(a.T @ u)
    sig = s.reshape(-1)
    res["residual_AtU_minus_VS_max_over_sigma1"] = float(((atu - vt.T * sig[None, :]).abs().max() / sig[0]).item())
    sig_host = sig.detach().cpu().numpy().copy()
    res["sigma_vs_golden_rel"] = None
    if write_golden and rk.rank == 0:
        np.savez(write_golden, sigma=sig_host, rows=rows, cols=n, k=k, q=q, p=p, seed=seed, workload=workload)
    try:
        g = np.load(GOLDEN_SIGMA)
        if (int(g["rows"]), int(g["cols"]), int(g["k"]), int(g["q"]), int(g["p"]), int(g["seed"])) == (rows, n, k, q, p, seed):
            res["sigma_vs_golden_rel"] = float(np.max(np.abs(sig_host - g["sigma"]) / np.abs(g["sigma"])))
            res["golden"] = "tests/golden/bench_c3_sigma.npz (1-GPU engine run of this seed; the engine is checked against the oracle in the small case below and in tests/)"
    except Exception:
        pass

    # a smaller case of the same shape class, sharded over the same ranks, against the CPU oracle (same A, same Omega)
    from oracle import ref_rsvd
    rows_s = 32768
    l = min(k + p, n)
    g = torch.Generator(device=a.device)
    g.manual_seed(seed * 1000 + 999)
    a_small = torch.randn((rows_s, n), dtype=torch.float64, device=a.device, generator=g)
    omega = np.random.default_rng(seed + 17).standard_normal((n, l))
    per = rows_s // rk.world
    mine = a_small[rk.rank * per:(rk.rank + 1) * per].contiguous()
    us, ss, vts = cb.rsvd(mine, k, q, p, omega=torch.from_numpy(omega).to(a.device), ctx=ctx, comm=comm,
                          global_rows=rows_s if comm is not None else None)
    u_all = rk.gather_rows(us.contiguous())
    small = {"rows": rows_s, "cols": n, "sharded_over": rk.world}
    if rk.rank == 0:
        u0, s0, vt0 = ref_rsvd.random_svd(a_small.cpu().numpy(), k, q, p, omega=omega)
        small["sigma_rel"] = ref_rsvd.sigma_rel_err(s0, ss.cpu().numpy())
        small["sin_u"] = ref_rsvd.subspace_sine(u0, u_all.cpu().numpy())
        small["sin_v"] = ref_rsvd.subspace_sine(vt0.T, vts.cpu().numpy().T)
    res["small_case_vs_oracle"] = small
    del a_small, mine, us, vts, u_all
    ok = (res["ortho_u_max"] < 1e-11 and res["ortho_v_max"] < 1e-11 and
          res["residual_AtU_minus_VS_max_over_sigma1"] < 1e-10 and
          (res["sigma_vs_golden_rel"] is None or res["sigma_vs_golden_rel"] < TOL_SIGMA))
    if rk.rank == 0:
        ok = ok and small["sigma_rel"] < TOL_SIGMA and small["sin_u"] < TOL_ANGLE and small["sin_v"] < TOL_ANGLE
    res["ok"] = bool(ok)
    return res


def other_config(torch, cb, device, name, peak64, hbm_peak, seed):
    """One compact line for a BASELINE config that is not the headline: device-resident time per call (CUDA events),
    whole-call fraction of its per-pass roofline, parity against the CPU oracle on the same A and Omega."""
    from oracle import ref_rsvd
    rows, n, k, q, p = WORKLOADS[name]
    l = min(k + p, n)
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000 + 77)
    a = torch.randn((rows, n), dtype=torch.float64, device=device, generator=g)
    omega = np.random.default_rng(seed + 3).standard_normal((n, l))
    om = torch.from_numpy(omega).to(device)
    ctx = cb.Context(device.index)
    reps = 20 if rows * n <= (1 << 22) else 5
    for _ in range(3):
        out = cb.rsvd(a, k, q, p, omega=om, ctx=ctx)
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        out = cb.rsvd(a, k, q, p, omega=om, ctx=ctx)
    e1.record()
    torch.cuda.synchronize(device)
    wall_ms = (time.perf_counter() - t0) * 1e3 / reps
    ms = float(e0.elapsed_time(e1)) / reps
    tm = cb.last_timings()
    passes = 2 + 2 * q
    t_flop = 2.0 * rows * n * l / (peak64 * 1e12) * 1e3
    t_hbm = rows * n * 8.0 / (hbm_peak * 1e9) * 1e3
    bound_ms = passes * max(t_flop, t_hbm)
    u0, s0, vt0 = ref_rsvd.random_svd(a.cpu().numpy(), k, q, p, omega=omega)
    u, s, vt = (x.cpu().numpy() for x in out)
    line = {"workload": f"{name}: {rows}x{n}, n_rank={k}, n_iters={q}, n_oversamples={p} (l={l}), device-resident",
            "ms_per_call": ms, "host_wall_ms_per_call": wall_ms, "calls_timed": reps,
            "gflops": flops_of(rows, n, l, q) / (ms * 1e-3) * 1e-9,
            "bound": "hbm" if t_hbm > t_flop else "tensor", "roofline_bound_ms": bound_ms,
            "whole_call_frac_of_per_pass_roofline": bound_ms / ms,
            "pass_kernel_avg_ms": (tm["pass_ms"] / tm["pass_launches"]) if tm["pass_launches"] else None,
            "pass_kernel_frac": (max(t_flop, t_hbm) / (tm["pass_ms"] / tm["pass_launches"])) if tm["pass_launches"] else None,
            "gpu_launches": tm["gpu_launches"], "fused_single_kernel": bool(tm.get("fused_small", 0)),
            "parity_vs_oracle": {"sigma_rel": ref_rsvd.sigma_rel_err(s0, s), "sin_u": ref_rsvd.subspace_sine(u0, u),
                                 "sin_v": ref_rsvd.subspace_sine(vt0.T, vt.T)}}
    pv = line["parity_vs_oracle"]
    line["parity_ok"] = bool(pv["sigma_rel"] < TOL_SIGMA and pv["sin_u"] < TOL_ANGLE and pv["sin_v"] < TOL_ANGLE)
    del a, out
    ctx.close()
    torch.cuda.empty_cache()
    return line


def run_ours(args):
    import torch
    import torch.distributed as dist
    import corrla_rs_b200 as cb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run (one process per GPU)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    numa = None
    if world > 1:
        # several ranks copy gigabytes at once: keep each rank's staging memory on the socket of its GPU
        from corrla_rs_b200 import hostnuma
        numa = hostnuma.bind_to_gpu_numa_node(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        comm = cb.ShardComm(device=local_rank)
    rk = Ranks(torch, dist, device, world, rank)

    rows, n, k, q, p = WORKLOADS[args.workload]
    if args.rows:
        rows = args.rows
    l = min(k + p, n)
    a, r0, r1 = make_shard(torch, device, rows, n, rank, world, args.seed)
    m_local = r1 - r0
    ctx = cb.Context(local_rank)

    def step_device(i):
        out = cb.rsvd(a, k, q, p, seed=args.seed + 100, ctx=ctx, comm=comm, global_rows=rows)
        return out, cb.last_timings()

    out = None
    for i in range(args.warmup):
        out, _ = step_device(i)
    del out
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    rk.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    pass_ms = 0.0
    pass_launches = 0
    pass_flops = 0.0
    t_wall0 = time.perf_counter()
    ev0.record()
    for i in range(args.steps):
        out, tm = step_device(i)
        launches += tm["gpu_launches"]
        pass_ms += tm["pass_ms"]; pass_launches += tm["pass_launches"]; pass_flops = tm["pass_flops"]
    ev1.record()
    rk.barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = rk.max(float(ev0.elapsed_time(ev1)))
    ms_per_step = ms_total / args.steps
    value = flops_of(rows, n, l, q) / (ms_per_step * 1e-3) * 1e-9
    s_dev = out[1].ravel()[:3].tolist()
    step_tm = dict(tm)

    # ---- self-check of the result of the last timed step (not timed)
    parity = None
    if not args.no_parity:
        parity = parity_block(torch, cb, rk, a, out, k, q, p, rows, n, args.seed, ctx, comm, args.workload,
                              args.write_golden)
        ok_all = rk.sum(1.0 if parity["ok"] else 0.0) == float(world)
        parity["ok_all_ranks"] = bool(ok_all)
    del out

    # ---- FP64 yardstick measured in this run (rank 0's GPU), next to the committed figure
    peak_file, peak_file_src = fp64_peak_file()
    peak_run = None
    if not args.no_peak:
        try:
            peak_run = measure_fp64_peak(torch, device) if rank == 0 else None
        except Exception as exc:
            peak_run = {"error": repr(exc)[:200]}
        rk.barrier()
    live = bool(peak_run and "sustained" in peak_run)
    peak = peak_run["sustained"] if live else peak_file
    peak_src = ("measured in this run: torch.matmul f64 8192^3 (cuBLAS DGEMM) back to back for 1 s on rank 0's GPU"
                if live else "file: " + peak_file_src)
    if world > 1:
        peak = rk.max(peak if (rank == 0 or args.no_peak) else 0.0)        # every rank uses rank 0's measurement

    # roofline of the dominant kernel (the DMMA GEMM streaming A): algorithmic flops per launch on this GPU divided by
    # the average CUDA-event duration of those launches inside the timed region; max over ranks of the duration.
    avg_pass_ms = rk.max(pass_ms / max(pass_launches, 1))
    hbm_peak = None
    try:
        hbm_peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0          # fallback stated in B200_PROFILING.md
    # a pass does 2*m*n*l flops on 8*m*n bytes: l/4 flop per byte against the machine balance decides the bound
    hbm_bound = (l / 4.0) < (peak * 1e12) / (hbm_peak * 1e9)
    traffic, traffic_src = measured_traffic(m_local, n)
    if hbm_bound:
        achieved = (m_local * n * 8.0) / (avg_pass_ms * 1e-3) * 1e-9 if pass_launches else None
        bound, unit, rpeak = "hbm", "GB/s", hbm_peak
        peak_src = "measured: MEASURED_PEAKS.json hbm_gbs" if (ROOT / "MEASURED_PEAKS.json").exists() else "fallback 6650 GB/s (B200_PROFILING.md)"
    else:
        achieved = pass_flops / (avg_pass_ms * 1e-3) * 1e-12 if pass_launches else None
        bound, unit, rpeak = "tensor", "TFLOP/s", peak
    roofline = {"bound": bound, "achieved": achieved, "peak": rpeak, "unit": unit,
                "frac": (achieved / rpeak) if achieved else None, "traffic": traffic, "traffic_source": traffic_src,
                "kernel": "skinny_gemm_kernel (DMMA.8x8x4 + TMA), one launch = one pass over this GPU's rows of A",
                "flops_per_launch": pass_flops, "avg_launch_ms": avg_pass_ms, "launches_timed": pass_launches,
                "hbm_bytes_per_launch_algorithmic": m_local * n * 8.0,
                "peak_source": peak_src,
                "peak_measured_in_run": peak_run, "peak_file": peak_file, "peak_file_source": peak_file_src,
                "peak_dmma_issue_rate": DMMA_ISSUE_PEAK,
                "frac_of_dmma_issue_rate": (achieved / DMMA_ISSUE_PEAK) if (achieved and not hbm_bound) else None,
                "frac_of_file_peak": (achieved / peak_file) if (achieved and not hbm_bound) else None,
                "whole_call_frac_of_per_pass_roofline": flops_of(rows, n, l, q) / world / (peak * 1e12) / (ms_per_step * 1e-3),
                "non_pass_ms_per_step": ms_per_step - avg_pass_ms * (pass_launches / max(args.steps, 1))}

    # ---- the other BASELINE configs, one compact entry each (single GPU only)
    other = None
    if world == 1 and not args.no_other_configs and args.workload == "c3":
        other = {}
        for name in ("c1", "c5", "c2"):
            try:
                other[name] = other_config(torch, cb, device, name, peak, hbm_peak, args.seed)
            except Exception as exc:
                other[name] = {"error": repr(exc)[:300]}

    # ---- end to end: host (pinned) buffers through the public API; H2D of A and D2H of U, S, Vt inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((m_local, n), dtype=torch.float64, pin_memory=True)
            host.copy_(a)
            torch.cuda.synchronize(device)
            a_host = host.numpy()
            del a
            torch.cuda.empty_cache()
            # caller-owned pinned outputs, reused by every call (out=): the results arrive by direct DMA, no per-call
            # multi-gigabyte allocation, no first-touch page faults
            u_h = torch.empty((k, m_local), dtype=torch.float64, pin_memory=True).numpy().T
            s_h = torch.empty((1, k), dtype=torch.float64, pin_memory=True).numpy().T
            vt_h = torch.empty((n, k), dtype=torch.float64, pin_memory=True).numpy().T
            outs = (u_h, s_h, vt_h)
            cb.rsvd(a_host, k, q, p, seed=args.seed + 100, ctx=ctx, comm=comm, global_rows=rows, out=outs)      # warm-up
            rk.barrier()
            t0 = time.perf_counter()
            for i in range(args.e2e_steps):
                u, s, vt = cb.rsvd(a_host, k, q, p, seed=args.seed + 100, ctx=ctx, comm=comm, global_rows=rows, out=outs)
            rk.barrier()
            dt = rk.max(time.perf_counter() - t0) / args.e2e_steps
            tm = cb.last_timings()
            h2d_bytes = int(rk.sum(m_local * n * 8))
            e2e = {"value": flops_of(rows, n, l, q) / dt * 1e-9, "unit": "GFLOP/s",
                   "h2d_bytes_per_step": h2d_bytes,
                   "d2h_bytes_per_step": int(rk.sum(m_local * k * 8) + (k + k * n) * 8),
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                   "breakdown_ms_rank0": {"h2d": tm["h2d_ms"], "device": tm["device_ms"], "d2h": tm["d2h_ms"]},
                   "h2d_gbs_rank0": m_local * n * 8 / max(tm["h2d_ms"], 1e-9) * 1e-6,
                   "h2d_gbs_aggregate": h2d_bytes / max(rk.max(tm["h2d_ms"]), 1e-9) * 1e-6,
                   "d2h_gbs_rank0": (m_local * k + k + k * n) * 8 / max(tm["d2h_ms"], 1e-9) * 1e-6,
                   "numa_binding_rank0": numa,
                   "sigma_head": np.asarray(s).ravel()[:3].tolist(),
                   "api": "corrla_rs_b200.rsvd(numpy array in pinned host memory, out=caller-owned pinned arrays) -> numpy arrays"}
        except Exception as exc:   # e.g. not enough pinned host memory on the box
            e2e = {"value": None, "unit": "GFLOP/s", "error": repr(exc)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = use_all_cpu_threads()
        sample = args.cpu_sample_rows or max(1024, min(rows, rows // 16))
        if sample < n:
            sample = rows          # a row slice of a (near-)square matrix would be a different (fat) problem: time it whole
        times = cpu_time_oracle(sample, n, k, q, p, args.seed, repeats=1, warmup=0)
        cms = 1e3 * times[0]
        cpu_baseline = {"value": flops_of(sample, n, l, q) / (cms * 1e-3) * 1e-9, "unit": "GFLOP/s", "cores": cores,
                        "kind": "port",
                        "sample": f"{sample} of {rows} rows x {n} cols ({'1/16 row slice' if sample * 16 == rows else 'row slice'}), "
                                  f"same k/q/p, numpy+OpenBLAS restatement of random_svd.rs (not faer); {cms:.0f} ms measured, "
                                  f"x{rows / sample:.0f} linear extrapolation = {cms * rows / sample:.0f} ms for the full matrix"}

    rc = 0
    if rank == 0:
        line = {
            "metric": "rsvd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, rows, n, k, q, p, world),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches, "sigma_head": s_dev, "parity": parity, "other_configs": other,
            "step_detail": {"jacobi_sweeps": step_tm.get("jacobi_sweeps"), "jacobi_converged": step_tm.get("jacobi_converged"),
                            "robust_qr_stages": step_tm.get("qr_third_passes"), "launches_per_call": step_tm.get("gpu_launches")},
            "collectives": ("none (single GPU)" if world == 1 else
                            (f"{step_tm['p2p_exchanges']} per call fused into the split-K reduction kernel over NVLink peer memory"
                             if step_tm["p2p_exchanges"] else "NCCL all-reduce")),
        }
        emit(line)
    if parity is not None and not parity.get("ok_all_ranks", True):
        sys.stderr.write(f"bench.py: PARITY FAILURE on rank {rank}: {json.dumps(parity)}\n")
        rc = 3
    if other:
        bad = [nm for nm, v in other.items() if isinstance(v, dict) and v.get("parity_ok") is False]
        if bad:
            sys.stderr.write(f"bench.py: PARITY FAILURE in other_configs: {bad}\n")
            rc = 3
    if world > 1:
        comm.close()
        dist.destroy_process_group()
    return rc


_REAL_STDOUT = None


def emit(line: dict):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else (NCCL banners, library chatter) was
    rerouted to stderr in main()."""
    out = os.fdopen(os.dup(_REAL_STDOUT), "w") if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    global _REAL_STDOUT
    # NCCL prints "NCCL version ..." on stdout from C; keep stdout clean for the single JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
