#!/usr/bin/env python
"""Average device time of a small RSVD (l = 110) where the fixed-cost kernels (Jacobi, Cholesky) dominate."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import corrla_rs_b200 as cb  # noqa: E402

g = torch.Generator(device="cuda")
g.manual_seed(1)
a = torch.randn((4096, 1024), dtype=torch.float64, device="cuda", generator=g)
for _ in range(3):
    cb.rsvd(a, 100, 4, 10, seed=3)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    cb.rsvd(a, 100, 4, 10, seed=3)
e1.record()
torch.cuda.synchronize()
print("ms per call", e0.elapsed_time(e1) / 20, "sweeps", cb.last_timings()["jacobi_sweeps"])
