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
region.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
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
    ap.add_argument("--cpu-sample-rows", type=int, default=0)
    ap.add_argument("--seed", type=int, default=5)
    return ap.parse_args()


def flops_of(m, n, l, q):
    return (2 + 2 * q) * 2.0 * m * n * l


def fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) holds only bf16 and HBM figures, so the FP64
    number is this repo's own measurement on the same pool: torch.matmul f64 8192^3 (cuBLAS DGEMM), same method as the
    bf16 entry -- profiles/r01_cublas_fp64_marks.json; the DMMA pipe itself peaks at 37.0 TFLOP/s
    (profiles/r01_fp64_pipe_microbench.txt)."""
    f = ROOT / "profiles" / "r01_cublas_fp64_marks.json"
    try:
        d = json.loads(f.read_text())
        return float(d["dgemm_8192_tflops_sustained"]), "measured: torch.matmul f64 8192^3 sustained (profiles/r01_cublas_fp64_marks.json)"
    except Exception:
        return 37.0, "fallback: DMMA.8x8x4 issue-rate peak at 1965 MHz"


def measured_traffic(rows_local, n):
    """dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, from an `ncu --set full`
    capture committed under profiles/ (bytes per row of A, scaled to this GPU's rows); None if no capture exists."""
    f = ROOT / "profiles" / "traffic.json"
    try:
        d = json.loads(f.read_text())
        if int(d["cols"]) != int(n):
            return None
        return float(d["dram_bytes_per_row"]) * rows_local
    except Exception:
        return None


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of random_svd.rs on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max((p.get("num_threads", 1) for p in threadpool_info()), default=1)
        return int(n)
    except Exception:
        return os.cpu_count() or 1


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
    cores = cpu_threads()
    line = {
        "impl": "reference",
        "metric": "rsvd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, rows, n, k, q, p, args.gpus),
        "cpu_baseline": {"value": value, "unit": "GFLOP/s", "cores": cores, "kind": "port",
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
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        comm = cb.ShardComm(device=local_rank)

    rows, n, k, q, p = WORKLOADS[args.workload]
    if args.rows:
        rows = args.rows
    l = min(k + p, n)
    a, r0, r1 = make_shard(torch, device, rows, n, rank, world, args.seed)
    m_local = r1 - r0
    ctx = cb.Context(local_rank)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

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
    barrier()
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
    barrier()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1)
    ms_total = max_over_ranks(float(ev0.elapsed_time(ev1)))
    ms_per_step = ms_total / args.steps
    value = flops_of(rows, n, l, q) / (ms_per_step * 1e-3) * 1e-9
    s_dev = out[1].ravel()[:3].tolist()

    # roofline of the dominant kernel (the DMMA GEMM streaming A): algorithmic flops per launch on this GPU divided by
    # the average CUDA-event duration of those launches inside the timed region; max over ranks of the duration.
    peak, peak_src = fp64_peak()
    avg_pass_ms = max_over_ranks(pass_ms / max(pass_launches, 1))
    hbm_peak = None
    try:
        hbm_peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0          # fallback stated in B200_PROFILING.md
    # a pass does 2*m*n*l flops on 8*m*n bytes: l/4 flop per byte against the machine balance decides the bound
    hbm_bound = (l / 4.0) < (peak * 1e12) / (hbm_peak * 1e9)
    if hbm_bound:
        achieved = (m_local * n * 8.0) / (avg_pass_ms * 1e-3) * 1e-9 if pass_launches else None
        bound, unit, rpeak = "hbm", "GB/s", hbm_peak
        peak_src = "measured: MEASURED_PEAKS.json hbm_gbs" if (ROOT / "MEASURED_PEAKS.json").exists() else "fallback 6650 GB/s (B200_PROFILING.md)"
    else:
        achieved = pass_flops / (avg_pass_ms * 1e-3) * 1e-12 if pass_launches else None
        bound, unit, rpeak = "tensor", "TFLOP/s", peak
    roofline = {"bound": bound, "achieved": achieved, "peak": rpeak, "unit": unit,
                "frac": (achieved / rpeak) if achieved else None, "traffic": measured_traffic(m_local, n),
                "kernel": "skinny_gemm_kernel (DMMA.8x8x4 + TMA), one launch = one pass over this GPU's rows of A",
                "flops_per_launch": pass_flops, "avg_launch_ms": avg_pass_ms, "launches_timed": pass_launches,
                "hbm_bytes_per_launch_algorithmic": m_local * n * 8.0,
                "peak_source": peak_src,
                "whole_call_frac_of_per_pass_roofline": flops_of(rows, n, l, q) / world / (peak * 1e12) / (ms_per_step * 1e-3)}
    del out

    # end to end: host (pinned) buffers through the public API; H2D of A and D2H of U, S, Vt inside the timed region
    e2e = None
    if not args.no_e2e:
        try:
            host = torch.empty((m_local, n), dtype=torch.float64, pin_memory=True)
            host.copy_(a)
            torch.cuda.synchronize(device)
            a_host = host.numpy()
            del a
            torch.cuda.empty_cache()
            cb.rsvd(a_host, k, q, p, seed=args.seed + 100, ctx=ctx, comm=comm, global_rows=rows)      # warm-up
            barrier()
            t0 = time.perf_counter()
            for i in range(args.e2e_steps):
                u, s, vt = cb.rsvd(a_host, k, q, p, seed=args.seed + 100, ctx=ctx, comm=comm, global_rows=rows)
            barrier()
            dt = max_over_ranks(time.perf_counter() - t0) / args.e2e_steps
            tm = cb.last_timings()
            e2e = {"value": flops_of(rows, n, l, q) / dt * 1e-9, "unit": "GFLOP/s",
                   "h2d_bytes_per_step": int(sum_over_ranks(m_local * n * 8)),
                   "d2h_bytes_per_step": int(sum_over_ranks(m_local * k * 8) + (k + k * n) * 8),
                   "ms_per_step": dt * 1e3, "steps": args.e2e_steps,
                   "breakdown_ms_rank0": {"h2d": tm["h2d_ms"], "device": tm["device_ms"], "d2h": tm["d2h_ms"]},
                   "sigma_head": np.asarray(s).ravel()[:3].tolist(),
                   "api": "corrla_rs_b200.rsvd(numpy array in pinned host memory) -> numpy arrays"}
        except Exception as exc:   # e.g. not enough pinned host memory on the box
            e2e = {"value": None, "unit": "GFLOP/s", "error": repr(exc)[:200], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = args.cpu_sample_rows or max(1024, min(rows, rows // 16))
        if sample < n:
            sample = rows          # a row slice of a (near-)square matrix would be a different (fat) problem: time it whole
        times = cpu_time_oracle(sample, n, k, q, p, args.seed, repeats=1, warmup=0)
        cms = 1e3 * times[0]
        cpu_baseline = {"value": flops_of(sample, n, l, q) / (cms * 1e-3) * 1e-9, "unit": "GFLOP/s", "cores": cpu_threads(),
                        "kind": "port",
                        "sample": f"{sample} of {rows} rows x {n} cols ({'1/16 row slice' if sample * 16 == rows else 'row slice'}), "
                                  f"same k/q/p, numpy+OpenBLAS restatement of random_svd.rs (not faer); {cms:.0f} ms measured, "
                                  f"x{rows / sample:.0f} linear extrapolation = {cms * rows / sample:.0f} ms for the full matrix"}

    if rank == 0:
        line = {
            "metric": "rsvd_gflops", "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.workload, rows, n, k, q, p, world),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "clocks": clocks,
            "gpu_launches": launches, "sigma_head": s_dev,
            "collectives": ("none (single GPU)" if world == 1 else
                            (f"{tm['p2p_exchanges']} per call fused into the split-K reduction kernel over NVLink peer memory"
                             if tm["p2p_exchanges"] else "NCCL all-reduce")),
        }
        emit(line)
    if world > 1:
        comm.close()
        dist.destroy_process_group()
    return 0


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
