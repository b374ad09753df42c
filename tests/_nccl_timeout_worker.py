"""Worker for tests/test_gpu_multi.py::test_peer_exchange_timeout_fails_on_every_rank (2 ranks, 2 GPUs)."""
import ctypes as C
import json
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def main():
    out_dir = Path(sys.argv[1])
    local = int(os.environ.get("LOCAL_RANK", "0"))
    rank = int(os.environ.get("RANK", "0"))
    if rank == 1:
        os.environ["CORRLA_B200_XCHG_TIMEOUT_CYCLES"] = "1"      # read when the communicator is created
    import corrla_rs_b200 as cb
    from corrla_rs_b200 import _ffi
    from oracle import ref_rsvd
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = cb.ShardComm(device=local)
    lib = _ffi.load()
    rng = np.random.default_rng(11)
    m, n, k, q, p = 8000, 128, 16, 3, 8
    a = rng.standard_normal((m, n))
    omega = rng.standard_normal((n, k + p))
    half = m // 2
    shard = np.ascontiguousarray(a[rank * half:(rank + 1) * half])
    # warm call so that both ranks have their buffers and kernels loaded (no timeouts yet on rank 0; rank 1 may already
    # time out here: that is the point, its status is collected below)
    ctx = cb.Context(local)
    u = np.zeros((half, k), order="F"); s = np.zeros(k); vt = np.zeros((k, n), order="F")
    o = _ffi.RsvdOpts(); lib.corrla_rsvd_opts_default(C.byref(o))
    o.seed = 1; o.ctx = ctx.handle; o.comm = comm.handle; o.global_rows = m
    dist.barrier()
    if rank == 0:
        time.sleep(1.0)                                           # rank 1 runs ahead and cannot see rank 0's epochs
    st = lib.corrla_rsvd_f64(shard.ctypes.data, half, n, n, 1, k, q, p, C.byref(o), u.ctypes.data, s.ctypes.data,
                             vt.ctypes.data, None)                # timings == NULL
    statuses = [None, None]
    dist.all_gather_object(statuses, int(st))
    p2p = cb.last_timings() is None                               # no Python-level call yet
    comm.close()
    # a fresh communicator (normal timeout) works again
    os.environ.pop("CORRLA_B200_XCHG_TIMEOUT_CYCLES", None)
    comm2 = cb.ShardComm(device=local)
    out = cb.rsvd(shard, k, q, p, omega=omega, comm=comm2, global_rows=m, ctx=ctx)
    used = cb.last_timings()["p2p_exchanges"] > 0
    after = [None, None]
    dist.all_gather_object(after, 0)
    if rank == 0:
        _, s0, _ = ref_rsvd.random_svd(a, k, q, p, omega=omega)
        res = {"status": statuses, "p2p": bool(used and p2p), "after_status": after,
               "after_sigma_rel": ref_rsvd.sigma_rel_err(s0, out[1])}
        (out_dir / "timeout.json").write_text(json.dumps(res))
    comm2.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
