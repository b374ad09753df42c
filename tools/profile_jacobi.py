#!/usr/bin/env python
"""One small RSVD with l = 110 (the Jacobi SVD of the 110 x 110 core runs once) for ncu -k regex:jacobi."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

import corrla_rs_b200 as cb  # noqa: E402

g = torch.Generator(device="cuda")
g.manual_seed(1)
a = torch.randn((20000, 1024), dtype=torch.float64, device="cuda", generator=g)
k = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for _ in range(2):
    u, s, vt = cb.rsvd(a, k, 4, 10, seed=3)
torch.cuda.synchronize()
print("ok", cb.last_timings()["jacobi_sweeps"])
