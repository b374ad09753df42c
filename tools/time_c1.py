"""Time BASELINE config C1 (100 x 100, n_rank 10, n_iters 12, n_oversamples 8) on the device-resident and host paths,
and print the fused kernel's phase breakdown (CORRLA_B200_FUSED_PROFILE=1)."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import corrla_rs_b200 as cb

rng = np.random.default_rng(1)
a = rng.standard_normal((100, 100))
omega = rng.standard_normal((100, 18))
ad, od = torch.from_numpy(a).cuda(), torch.from_numpy(omega).cuda()
ctx = cb.Context(0)
for name, x, om in (("device", ad, od), ("host", a, omega)):
    for _ in range(5):
        cb.rsvd(x, 10, 12, 8, omega=om, ctx=ctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev = []
    for _ in range(50):
        cb.rsvd(x, 10, 12, 8, omega=om, ctx=ctx)
        dev.append(cb.last_timings()["device_ms"])
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / 50 * 1e3
    print(f"C1 {name}: wall {wall:.3f} ms per call, kernel (CUDA events) median {np.median(dev):.3f} ms, fused={cb.last_timings()['fused_small']}")
os.environ["CORRLA_B200_FUSED_PROFILE"] = "1"
cb.rsvd(ad, 10, 12, 8, omega=od, ctx=ctx)
torch.cuda.synchronize()
